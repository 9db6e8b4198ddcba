#!/bin/bash
# A/B of build variants of the same library (ICIKT_LIB_PATH), quick benches
mkdir -p gpurun_out
export ICIKT_REQUIRE_GPU=1
ICIKT_LIB_PATH=$PWD/icikendalltau_b200/variant_v2rg.so timeout 900 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -2
ICIKT_LIB_PATH=$PWD/icikendalltau_b200/variant_v2rg.so ICIKT_FUZZ_SIZES=24576,28672,32768,40000,57344,64512,64513,65535 timeout 300 python tools/fuzz.py 100 779 2>&1 | tail -2
ICIKT_LIB_PATH=$PWD/icikendalltau_b200/variant_v2rg.so ICIKT_FORCE_GMEM=1 ICIKT_FUZZ_SIZES=33,257,1000,2049,5000,8193,10000 timeout 300 python tools/fuzz.py 50 780 2>&1 | tail -2
run() { ICIKT_LIB_PATH=$PWD/icikendalltau_b200/variant_$1.so timeout 600 python bench.py --workload $2 --steps 5 --warmup 3 --quick 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1 $2', round(d['value']), 'k2', round(r['k2_ms'],3), 'frac', round(r['frac'],3))"; }
{
for v in base v2rg v2all base v2rg; do run $v config4; done
for v in base v2all base v2all; do run $v config1; run $v target; done
} | tee gpurun_out/r02_variants.txt
