#!/bin/bash
mkdir -p gpurun_out
run() { ICIKT_LIB_PATH=$PWD/icikendalltau_b200/$1.so timeout 600 python bench.py --workload $2 --steps 5 --warmup 3 --quick 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1 $2', round(d['value']), 'k2', round(r['k2_ms'],3), 'frac', round(r['frac'],3))"; }
{
for v in variant_v1 variant_v4 variant_v7; do run $v target; run $v config1; run $v config5; done
} | tee gpurun_out/r02_variants.txt
