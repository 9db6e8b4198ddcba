#!/bin/bash
# closing run of the round: instruction / DRAM counts of the final pair kernel first (bench.py's issue roofline reads
# them), then the bench lines of the four tie-free workloads
mkdir -p gpurun_out
for wl in target config5 config3 config2; do
timeout 600 ncu --metrics smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:pairs_tiled -c 3 --csv --log-file gpurun_out/r02w_counts_$wl.csv python bench.py --workload $wl --steps 1 --warmup 3 --quick > gpurun_out/r02w_counts_$wl.log 2>&1; echo $wl ncu_exit=$?
done
python tools/k2_counts_json.py gpurun_out/r02w_counts_ "round 2 final K2"
timeout 600 python bench.py > gpurun_out/r02z_bench_target.json 2> gpurun_out/r02z_bench_target.err; echo bench_exit=$?
for wl in config5 config3 config2; do
timeout 600 python bench.py --workload $wl --cpu-budget 5 > gpurun_out/r02z_bench_$wl.json 2> gpurun_out/r02z_bench_$wl.err; echo $wl exit=$?
done
for wl in target config5 config3 config2; do
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02z_bench_$wl.json").read().strip().splitlines()[-1])
    r=d["roofline"]; print("$wl", round(d["value"]), "pairs/s ms/step", round(d["ms_per_step"],3), "k1", round(r["k1_ms"],3), "k2", round(r["k2_ms"],3), "frac", round(r["frac"],3), "issue", r.get("issue") and round(r["issue"]["frac"],3), "e2e", round(d["e2e"]["value"]), "pageable", round(d["e2e_pageable"]["value"]), "cpu", round(d["cpu_baseline"]["value"]), "parity", d["parity_sample"]["ok"])
except Exception as e: print("$wl", "no line", e)
PY
done
