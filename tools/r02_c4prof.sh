#!/bin/bash
# config 4: full ncu capture of the in-place pair kernel (the second matching launch: the first is tier 0's early exit)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:pairs_tiled --launch-skip 1 -c 1 -o gpurun_out/r02u_k2_config4 -f \
  python bench.py --workload config4 --steps 1 --warmup 1 --quick --cols 60 > gpurun_out/r02u_ncu.log 2>&1
ls -la gpurun_out/r02u*
