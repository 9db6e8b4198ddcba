#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pairs_tiled -c 3 -o gpurun_out/r02h_k2_config4 -f python bench.py --workload config4 --steps 1 --warmup 3 --quick --cols 60 > gpurun_out/r02h_ncu_k2c4.log 2>&1; echo ncu_exit=$?
timeout 600 python bench.py --workload config4 --steps 5 --warmup 3 --quick --cols 60 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('config4 x60', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'k1', round(r['k1_ms'],3), 'k2', round(r['k2_ms'],3))"
