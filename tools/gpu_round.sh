set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$? | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
bash tools/sweep.sh
prof() { # name workload cols
ncu --set full --clock-control none --import-source on -k regex:pairs_tiled --launch-skip 3 --launch-count 1 -f -o gpurun_out/prof_k2v6_$1 python bench.py --workload $2 --cols $3 --steps 1 --warmup 3 --quick > gpurun_out/ncu_$1.log 2>&1
}
prof target target 300
