set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$? | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python tools/gpu_diag.py > gpurun_out/diag.log 2>&1; tail -2 gpurun_out/diag.log
python bench.py --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/q_config2.json
for wl in config5 target; do
python bench.py --workload $wl --steps 3 --warmup 3 --quick 2>/dev/null | tail -1 > gpurun_out/q_$wl.json
done
prof() { # name workload cols
ncu --set full --clock-control none --import-source on -k regex:pairs_tiled --launch-skip 6 --launch-count 1 -f -o gpurun_out/prof_k2v7_$1 python bench.py --workload $2 --cols $3 --steps 1 --warmup 3 --quick > gpurun_out/ncu_$1.log 2>&1
}
prof target target 300
