#!/bin/bash
# A/B of build variants of the same library (ICIKT_LIB_PATH), quick benches
mkdir -p gpurun_out
run() { ICIKT_LIB_PATH=$PWD/icikendalltau_b200/variant_$1.so timeout 600 python bench.py --workload $2 --steps 5 --warmup 3 --quick 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1 $2', round(d['value']), 'k2', round(r['k2_ms'],3), 'frac', round(r['frac'],3))"; }
{
for v in base seq1 seq1p seq2 base seq1 seq1p; do run $v config4; done
for v in base seq2 base seq2; do run $v config1; run $v target; done
} | tee gpurun_out/r02_variants.txt
