#!/bin/bash
# refresh of the 8-GPU evidence with the final code: strong-scaled target at N = 8, 4, 2; matrix output; one-call multi
mkdir -p gpurun_out
for N in 8 4 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02q_bench_target_n$N.json 2> gpurun_out/r02q_bench_target_n$N.err; echo bench$N exit=$?
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02q_bench_target_n$N.json").read().strip().splitlines()[-1])
    r=d["roofline"]; print("N=$N", round(d["value"]), "pairs/s  ms/step", round(d["ms_per_step"],3), "k1", round(r["k1_ms"],3), "k2", round(r["k2_ms"],3), "e2e", round(d["e2e"]["value"]), "e2e_pageable", round(d["e2e_pageable"]["value"]), "parity", d.get("parity_sample",{}).get("ok"))
except Exception as e: print("N=$N no line", e)
PY
done
timeout 900 python tools/matrix_multi_timing.py > gpurun_out/r02q_matrix_multi.txt 2>&1; tail -5 gpurun_out/r02q_matrix_multi.txt
timeout 600 python tools/multi_call_timing.py > gpurun_out/r02q_multi_call.txt 2>&1; tail -4 gpurun_out/r02q_multi_call.txt
