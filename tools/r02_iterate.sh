#!/bin/bash
export ICIKT_REQUIRE_GPU=1
mkdir -p gpurun_out
T=${1:-r02g}
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${T}_pytest_gpu.log 2>&1; echo pytest_exit=$? >> gpurun_out/${T}_pytest_gpu.log
tail -8 gpurun_out/${T}_pytest_gpu.log
for wl in target config2 config5 config4 config1; do
timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --quick > gpurun_out/${T}_q_$wl.json 2> gpurun_out/${T}_q_$wl.err; echo $wl exit=$?
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_q_$wl.json").read().strip().splitlines()[-1])
    r=d["roofline"]; print("$wl", round(d["value"]), "pairs/s  ms/step", round(d["ms_per_step"],3), "k1", round(r["k1_ms"],3), "k2", round(r["k2_ms"],3), "frac", round(r["frac"],3))
except Exception as e: print("$wl", "no line", e)
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pairs_tiled -c 1 -o gpurun_out/${T}_k2_target300 -f python bench.py --steps 1 --warmup 3 --quick --cols 300 > gpurun_out/${T}_ncu_k2.log 2>&1; echo ncu_k2_exit=$?
