set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$? | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for r in 5000 20000 40000; do
python bench.py --workload config4 --rows $r --cols 300 --steps 3 --warmup 3 --quick 2>/dev/null | tail -1 > gpurun_out/q_counts_$r.json
done
python bench.py --workload config1 --steps 10 --warmup 3 --quick 2>/dev/null | tail -1 > gpurun_out/q_config1.json
for W in 2 3 4 6 9; do ICIKT_WARPS=$W python bench.py --workload config1 --steps 10 --warmup 3 --quick 2>/dev/null | tail -1 > gpurun_out/q_config1_w$W.json; done
