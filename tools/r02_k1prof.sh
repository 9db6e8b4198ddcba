#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:column_sort --launch-skip 1 -c 1 -o gpurun_out/r02v_k1sort_config4 -f \
  python bench.py --workload config4 --steps 1 --warmup 1 --quick > gpurun_out/r02v_ncu.log 2>&1
ls -la gpurun_out/r02v*
