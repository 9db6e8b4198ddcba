#!/bin/bash
# Quick GPU check used while iterating: parity suite, then short benches of the main workloads.
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$? | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/q_config2.json
for wl in config1 config5 target; do
timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 --quick 2>/dev/null | tail -1 > gpurun_out/q_$wl.json
done
