set -x
python bench.py --workload config1 --steps 10 --warmup 3 --cpu-budget 5 2>/dev/null | tail -1 > gpurun_out/q_config1.json
