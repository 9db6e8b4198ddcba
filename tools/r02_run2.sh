#!/bin/bash
# round 2, GPU call 2: own radix sort in K1 -- parity suite, quick benches of every workload,
# launch list + full ncu capture of K2 (300 samples of the target shape) and of the column kernels.
export ICIKT_REQUIRE_GPU=1
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02b_pytest_gpu.log 2>&1; echo pytest_exit=$? >> gpurun_out/r02b_pytest_gpu.log
tail -15 gpurun_out/r02b_pytest_gpu.log
for wl in target config2 config5 config4 config1 config3; do
timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --quick > gpurun_out/r02b_q_$wl.json 2> gpurun_out/r02b_q_$wl.err; echo $wl exit=$?
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02b_q_$wl.json").read().strip().splitlines()[-1])
    r=d["roofline"]; print("$wl", round(d["value"]), "pairs/s  ms/step", round(d["ms_per_step"],3), "k1", round(r["k1_ms"],3), "k2", round(r["k2_ms"],3), "frac", round(r["frac"],3), "issue", r.get("issue") and round(r["issue"]["frac"],3), r.get("issue") and r["issue"]["peak_source"])
except Exception as e: print("$wl", "no line", e)
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02b_launches_target.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/r02b_ncu_launches.log 2>&1; echo ncu_exit=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pairs_tiled -c 1 -o gpurun_out/r02b_k2_target300 -f python bench.py --steps 1 --warmup 3 --quick --cols 300 > gpurun_out/r02b_ncu_k2.log 2>&1; echo ncu_k2_exit=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:column_ -c 3 -o gpurun_out/r02b_k1_target300 -f python bench.py --steps 1 --warmup 3 --quick --cols 300 > gpurun_out/r02b_ncu_k1.log 2>&1; echo ncu_k1_exit=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:column_fused -c 1 -o gpurun_out/r02b_k1_config2 -f python bench.py --workload config2 --steps 1 --warmup 3 --quick > gpurun_out/r02b_ncu_k1c2.log 2>&1; echo ncu_k1c2_exit=$?
