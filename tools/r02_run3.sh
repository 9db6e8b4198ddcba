#!/bin/bash
# round 2, GPU call 3 (2 GPUs): radix pass v2 + multi-GPU matrices -- parity suite, K1 timings
export ICIKT_REQUIRE_GPU=1
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02c_pytest_gpu.log 2>&1; echo pytest_exit=$? >> gpurun_out/r02c_pytest_gpu.log
tail -25 gpurun_out/r02c_pytest_gpu.log
for wl in target config2 config5 config4; do
timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --quick > gpurun_out/r02c_q_$wl.json 2> gpurun_out/r02c_q_$wl.err; echo $wl exit=$?
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c_q_$wl.json").read().strip().splitlines()[-1])
    r=d["roofline"]; print("$wl", round(d["value"]), "pairs/s  ms/step", round(d["ms_per_step"],3), "k1", round(r["k1_ms"],3), "k2", round(r["k2_ms"],3), "frac", round(r["frac"],3))
except Exception as e: print("$wl", "no line", e)
PY
done
python tools/matrix_multi_timing.py > gpurun_out/r02c_matrix_multi.txt 2>&1; cat gpurun_out/r02c_matrix_multi.txt
