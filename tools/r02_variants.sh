#!/bin/bash
# parity of the current build on the long-vector paths, then A/B against variant_base.so (ICIKT_LIB_PATH), quick benches
mkdir -p gpurun_out
export ICIKT_REQUIRE_GPU=1
timeout 900 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -2
ICIKT_FUZZ_SIZES=22529,24576,28672,32768,40000,41000,57344,64512,64513,65535 timeout 300 python tools/fuzz.py 120 781 2>&1 | tail -2
run() { ICIKT_LIB_PATH=$PWD/icikendalltau_b200/$1.so timeout 600 python bench.py --workload $2 --steps 5 --warmup 3 --quick $3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1 $2 $3', round(d['value']), 'k1', round(r['k1_ms'],3), 'k2', round(r['k2_ms'],3), 'frac', round(r['frac'],3))"; }
{
for v in variant_base libicikt_b200 variant_base libicikt_b200; do run $v config4; done
run libicikt_b200 target
run variant_base config3 "--rows 30000 --cols 300"
run libicikt_b200 config3 "--rows 30000 --cols 300"
} | tee gpurun_out/r02_variants.txt
