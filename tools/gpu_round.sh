set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$? | tee -a gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
python bench.py --workload config4 --steps 2 --warmup 3 --quick 2>/dev/null | tail -1 > gpurun_out/q_counts_60000.json
python bench.py --workload config4 --rows 50000 --cols 200 --steps 2 --warmup 3 --quick 2>/dev/null | tail -1 > gpurun_out/q_counts_50000.json
python bench.py --workload target --rows 60000 --cols 300 --steps 2 --warmup 3 --quick 2>/dev/null | tail -1 > gpurun_out/q_cont_60000.json
ICIKT_FORCE_GMEM=1 python bench.py --workload target --rows 60000 --cols 300 --steps 2 --warmup 3 --quick 2>/dev/null | tail -1 > gpurun_out/q_cont_60000_gmem.json
