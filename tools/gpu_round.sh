set -x
prof() { # name workload cols
ncu --set full --clock-control none --import-source on -k regex:pairs_tiled --launch-skip 6 --launch-count 1 -f -o gpurun_out/prof_r01_k2_$1 python bench.py --workload $2 --cols $3 --steps 1 --warmup 3 --quick > gpurun_out/ncu_$1.log 2>&1
}
prof config2 config2 100
prof target target 300
