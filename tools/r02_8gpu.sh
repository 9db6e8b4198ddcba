#!/bin/bash
# round 2, 8-GPU call: strong-scaled target bench at N = 8 and 4 (torchrun, NCCL table exchange), the
# multi-GPU tests with real peers, ici_kendalltau matrix output at config 5 on 1/2/4/8 GPUs, and the
# one-call icikt_all_pairs_multi timing.
export ICIKT_REQUIRE_GPU=1
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02i_bench_target_n$N.json 2> gpurun_out/r02i_bench_target_n$N.err; echo bench$N exit=$?
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02i_bench_target_n$N.json").read().strip().splitlines()[-1])
    r=d["roofline"]; print("N=$N", round(d["value"]), "pairs/s  ms/step", round(d["ms_per_step"],3), "k1", round(r["k1_ms"],3), "k2", round(r["k2_ms"],3), "e2e", round(d["e2e"]["value"]), "e2e_pageable", round(d["e2e_pageable"]["value"]), "parity", d.get("parity_sample",{}).get("ok"))
except Exception as e: print("N=$N no line", e)
PY
tail -3 gpurun_out/r02i_bench_target_n$N.err
done
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "multi or shard or target_full or shim" > gpurun_out/r02i_pytest_multi.log 2>&1; echo pytest_exit=$?; tail -5 gpurun_out/r02i_pytest_multi.log
timeout 900 python tools/matrix_multi_timing.py > gpurun_out/r02i_matrix_multi.txt 2>&1; cat gpurun_out/r02i_matrix_multi.txt | tail -6
timeout 600 python tools/multi_call_timing.py > gpurun_out/r02i_multi_call.txt 2>&1; tail -8 gpurun_out/r02i_multi_call.txt
