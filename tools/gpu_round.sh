set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$? | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for seed in 1 2 3 4 5 6 7 8; do ICIKT_TEST_SEED=$seed python -m pytest tests/test_gpu_parity.py -q -k "random_small or repeated_runs or uneven" 2>&1 | tail -1; done
