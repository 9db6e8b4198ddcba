set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$? | tee -a gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
python tools/gpu_diag.py > gpurun_out/diag.log 2>&1; tail -2 gpurun_out/diag.log
for r in 5000 20000 40000; do
python bench.py --workload config4 --rows $r --cols 300 --steps 3 --warmup 3 --quick 2>/dev/null | tail -1 > gpurun_out/q_counts_$r.json
done
python bench.py --workload config4 --steps 2 --warmup 3 --quick 2>/dev/null | tail -1 > gpurun_out/q_counts_60000.json
python bench.py --workload config1 --steps 10 --warmup 3 --quick 2>/dev/null | tail -1 > gpurun_out/q_config1.json
python bench.py --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/q_config2.json
python bench.py --workload target --steps 3 --warmup 3 --quick 2>/dev/null | tail -1 > gpurun_out/q_target.json
