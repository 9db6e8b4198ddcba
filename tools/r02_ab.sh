#!/bin/bash
mkdir -p gpurun_out
for wl in target config5 config2 config4; do
timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --quick 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$wl', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'k1', round(r['k1_ms'],3), 'k2', round(r['k2_ms'],3), 'frac', round(r['frac'],3))"
done
