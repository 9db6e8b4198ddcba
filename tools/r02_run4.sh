#!/bin/bash
export ICIKT_REQUIRE_GPU=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "multi_gpu" > gpurun_out/r02d_multi.log 2>&1; tail -30 gpurun_out/r02d_multi.log
timeout 900 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02d_pytest_gpu.log 2>&1; echo pytest_exit=$? >> gpurun_out/r02d_pytest_gpu.log
tail -25 gpurun_out/r02d_pytest_gpu.log
