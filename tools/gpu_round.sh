set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$? | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/q_config2.json
timeout 300 python bench.py --workload config5 --steps 3 --warmup 3 --quick 2>/dev/null | tail -1 > gpurun_out/q_config5.json
timeout 300 python bench.py --workload target --steps 3 --warmup 3 --quick 2>/dev/null | tail -1 > gpurun_out/q_target.json
timeout 300 python bench.py --workload config1 --steps 10 --warmup 3 --quick 2>/dev/null | tail -1 > gpurun_out/q_config1.json
