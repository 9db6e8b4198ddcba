#!/bin/bash
mkdir -p gpurun_out
export ICIKT_REQUIRE_GPU=1
timeout 900 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -2
timeout 300 python tools/fuzz.py 50 788 2>&1 | tail -2
run() { ICIKT_LIB_PATH=$PWD/icikendalltau_b200/$1.so timeout 600 python bench.py --workload $2 --steps 5 --warmup 3 --quick 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1 $2', round(d['value']), 'k2', round(r['k2_ms'],3), 'frac', round(r['frac'],3))"; }
{
for w in target config5 config2; do run variant_base $w; run libicikt_b200 $w; done
} | tee gpurun_out/r02_variants.txt
