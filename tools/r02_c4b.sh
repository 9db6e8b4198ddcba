#!/bin/bash
export ICIKT_REQUIRE_GPU=1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -2
ICIKT_FUZZ_SIZES=24576,28672,32768,40000,57344,64512,64513,65535 timeout 300 python tools/fuzz.py 150 777 > gpurun_out/r02t_fuzz_long.txt 2>&1; tail -2 gpurun_out/r02t_fuzz_long.txt
ICIKT_FORCE_GMEM=1 ICIKT_FUZZ_SIZES=33,257,1000,2049,5000,8193,10000 timeout 300 python tools/fuzz.py 60 778 > gpurun_out/r02t_fuzz_gmem.txt 2>&1; tail -2 gpurun_out/r02t_fuzz_gmem.txt
run() { ICIKT_LIB_PATH=$1 timeout 600 python bench.py --workload $2 --steps 5 --warmup 3 --quick 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$3 $2', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'k1', round(r['k1_ms'],3), 'k2', round(r['k2_ms'],3), 'frac', round(r['frac'],3))"; }
for i in 1 2; do
run $PWD/icikendalltau_b200/variant_base.so config4 base
run $PWD/icikendalltau_b200/libicikt_b200.so config4 new
done
run $PWD/icikendalltau_b200/libicikt_b200.so target new
