#!/bin/bash
# A/B of build variants of the same library (ICIKT_LIB_PATH), target + config 5, quick benches
mkdir -p gpurun_out
for v in base khm base khm; do
for wl in target config5; do
ICIKT_LIB_PATH=$PWD/icikendalltau_b200/variant_$v.so timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --quick 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$v $wl', round(d['value']), 'k2', round(r['k2_ms'],3), 'frac', round(r['frac'],3))"
done
done | tee gpurun_out/r02_variants.txt
