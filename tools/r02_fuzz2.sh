#!/bin/bash
export ICIKT_REQUIRE_GPU=1
mkdir -p gpurun_out
timeout 400 python tools/fuzz.py 280 20261018 > gpurun_out/r02n_fuzz_a.txt 2>&1; tail -2 gpurun_out/r02n_fuzz_a.txt
timeout 400 python tools/fuzz.py 280 4242 > gpurun_out/r02n_fuzz_b.txt 2>&1; tail -2 gpurun_out/r02n_fuzz_b.txt
for i in 1 2 3 4 5 6; do timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_gpu or repeated_runs" 2>&1 | tail -1; done
