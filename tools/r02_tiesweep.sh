#!/bin/bash
# config 4: thresholds between direct comparison and sorting of tie groups, with the round's in-place kernel
mkdir -p gpurun_out
run() { ICIKT_LARGE_TIE=$1 ICIKT_DIRECT_BUDGET=$2 timeout 300 python bench.py --workload config4 --steps 4 --warmup 2 --quick 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('large_tie $1 budget $2', round(d['value']), 'k1', round(r['k1_ms'],3), 'k2', round(r['k2_ms'],3), 'frac', round(r['frac'],3))"; }
{
for lt in 16 32 64 128 256; do run $lt 24; done
for b in 4 8 48; do run 128 $b; run 48 $b; done
} | tee gpurun_out/r02_tiesweep_config4.txt
