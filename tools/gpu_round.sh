set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$? | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python tools/gpu_diag.py > gpurun_out/diag.log 2>&1; tail -1 gpurun_out/diag.log
