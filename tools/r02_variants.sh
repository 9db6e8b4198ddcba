#!/bin/bash
# parity of the current build on the long-vector paths, then quick benches
mkdir -p gpurun_out
export ICIKT_REQUIRE_GPU=1
timeout 900 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -2
ICIKT_FUZZ_SIZES=22529,24576,28672,32768,40000,41000,57344,64512,64513,65535 timeout 300 python tools/fuzz.py 80 785 2>&1 | tail -2
ICIKT_SORT_GLOBAL=1 ICIKT_FUZZ_SIZES=8193,9000,10000,16384,20000,22528 timeout 300 python tools/fuzz.py 40 786 2>&1 | tail -2
run() { timeout 600 python bench.py --workload $1 --steps 5 --warmup 3 --quick $2 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1 $2', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'k1', round(r['k1_ms'],3), 'k2', round(r['k2_ms'],3), 'frac', round(r['frac'],3))"; }
{
run config4; run config4
run config3 "--rows 30000 --cols 300"
run config3 "--rows 60000 --cols 400"
} | tee gpurun_out/r02_variants.txt
