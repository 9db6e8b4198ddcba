#!/bin/bash
# Quick GPU check used while iterating: parity suite, then short benches of the main workloads.
set -x
export ICIKT_REQUIRE_GPU=1  # a silent skip of the -m gpu tests on the GPU box would read as green
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$? | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py 2>/dev/null | tail -1 > gpurun_out/q_target_default.json
for wl in config1 config2 config5; do
timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 --quick 2>/dev/null | tail -1 > gpurun_out/q_$wl.json
done
