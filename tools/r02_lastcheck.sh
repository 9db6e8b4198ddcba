#!/bin/bash
export ICIKT_REQUIRE_GPU=1
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bitmap or pipelined or random_parity" 2>&1 | tail -3
