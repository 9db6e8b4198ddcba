#!/bin/bash
# round 2, final evidence on one GPU: smoke, parity suite, fuzz campaign, the default bench line (target), the
# other workloads in full, the reference arm, the ncu launch list of the default command.
export ICIKT_REQUIRE_GPU=1
mkdir -p gpurun_out
T=r02z
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1; echo smoke_exit=$?; tail -1 gpurun_out/${T}_smoke.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${T}_pytest_gpu.log 2>&1; echo pytest_exit=$? >> gpurun_out/${T}_pytest_gpu.log; tail -4 gpurun_out/${T}_pytest_gpu.log
timeout 400 python tools/fuzz.py 80 777 > gpurun_out/${T}_fuzz.txt 2>&1; tail -3 gpurun_out/${T}_fuzz.txt
timeout 900 python bench.py > gpurun_out/${T}_bench_target.json 2> gpurun_out/${T}_bench_target.err; echo bench_exit=$?
for wl in config1 config2 config3 config4 config5; do
timeout 900 python bench.py --workload $wl --cpu-budget 8 > gpurun_out/${T}_bench_$wl.json 2> gpurun_out/${T}_bench_$wl.err; echo $wl exit=$?
done
for wl in target config1 config2 config3 config4 config5; do
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_bench_$wl.json").read().strip().splitlines()[-1])
    r=d["roofline"]; print("$wl", round(d["value"]), "pairs/s ms/step", round(d["ms_per_step"],3), "k1", round(r["k1_ms"],3), "k2", round(r["k2_ms"],3), "frac", round(r["frac"],3), "issue", r.get("issue") and round(r["issue"]["frac"],3), "e2e", round(d["e2e"]["value"]), "pageable", round(d["e2e_pageable"]["value"]), "cpu", round(d["cpu_baseline"]["value"]), d["cpu_baseline"]["cores"], "parity", d["parity_sample"]["ok"])
except Exception as e: print("$wl", "no line", e)
PY
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>/dev/null; tail -c 600 gpurun_out/${T}_bench_reference.json
timeout 600 python bench.py --kernel naive --workload config2 --no-cpu-baseline --steps 3 > gpurun_out/${T}_bench_naive_config2.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/${T}_bench_naive_config2.json').read().strip().splitlines()[-1]); print('naive config2', round(d['value']), 'pairs/s frac', round(d['roofline']['frac'],4))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_target.csv python bench.py --steps 2 --warmup 3 > gpurun_out/${T}_ncu_launches.log 2>&1; echo ncu_exit=$?
