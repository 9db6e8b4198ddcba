set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$? | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --workload config1 --steps 10 --warmup 3 --quick 2>/dev/null | tail -1 > gpurun_out/q_config1.json
timeout 300 python bench.py --workload config4 --rows 20000 --cols 300 --steps 3 --warmup 3 --quick 2>/dev/null | tail -1 > gpurun_out/q_counts_20000.json
timeout 300 python bench.py --workload config4 --rows 5000 --cols 300 --steps 3 --warmup 3 --quick 2>/dev/null | tail -1 > gpurun_out/q_counts_5000.json
timeout 300 python bench.py --workload config4 --steps 2 --warmup 3 --quick 2>/dev/null | tail -1 > gpurun_out/q_counts_60000.json
