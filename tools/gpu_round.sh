set -x
ncu --set full --clock-control none --import-source on -k regex:pairs_tiled --launch-skip 2 --launch-count 1 -f -o gpurun_out/prof_k2_config4 python bench.py --workload config4 --cols 40 --steps 1 --warmup 3 --quick > gpurun_out/ncu_config4.log 2>&1
tail -3 gpurun_out/ncu_config4.log
