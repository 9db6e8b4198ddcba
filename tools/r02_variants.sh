#!/bin/bash
mkdir -p gpurun_out
run() { ICIKT_LIB_PATH=$PWD/icikendalltau_b200/$1.so timeout 200 python bench.py --workload $2 --steps 4 --warmup 2 --quick 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1 $2', round(d['value']), 'k2', round(r['k2_ms'],3), 'frac', round(r['frac'],3))"; }
{ for v in libicikt_b200 variant_wide256 variant_list8 variant_list32; do run $v config4; done; } | tee gpurun_out/r02_variants.txt
