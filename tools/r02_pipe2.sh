#!/bin/bash
export ICIKT_REQUIRE_GPU=1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -2
for b in 8 16 32; do
  ICIKT_PIPELINE_BLOCKS=$b python tools/pipe_timing.py config5 2>&1 | grep "pipe " | sed "s/^/blocks=$b /"
  ICIKT_PIPELINE_BLOCKS=$b python tools/pipe_timing.py target 2>&1 | grep "pipe " | sed "s/^/blocks=$b /"
done
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02s_bench_target.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02s_bench_target.json').read().strip().splitlines()[-1]); r=d['roofline']
print('target', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'k2', round(r['k2_ms'],2), 'e2e', d['e2e'], 'pageable', round(d['e2e_pageable']['value']), 'parity', d['parity_sample']['ok'])"
