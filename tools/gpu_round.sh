set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$? | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python tools/gpu_diag.py > gpurun_out/diag.log 2>&1; tail -3 gpurun_out/diag.log
for wl in config2 config5 target; do
python bench.py --workload $wl --steps 5 --warmup 3 --quick 2>/dev/null | tail -1 > gpurun_out/q_$wl.json
done
ICIKT_NO_FUSED_COLUMNS=1 python bench.py --workload config2 --steps 5 --warmup 3 --quick 2>/dev/null | tail -1 > gpurun_out/q_config2_nofuse.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_config2.csv python bench.py --workload config2 --steps 2 --warmup 3 --quick > gpurun_out/ncu_launch.log 2>&1
