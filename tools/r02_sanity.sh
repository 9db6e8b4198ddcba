#!/bin/bash
export ICIKT_REQUIRE_GPU=1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02r_bench_target.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02r_bench_target.json').read().strip().splitlines()[-1]); r=d['roofline']
print('target', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'k1', round(r['k1_ms'],2), 'k2', round(r['k2_ms'],2), 'frac', round(r['frac'],3), 'issue', round(r['issue']['frac'],3), 'traffic', r['traffic'], 'e2e', round(d['e2e']['value']), 'pageable', round(d['e2e_pageable']['value']), 'parity', d['parity_sample']['ok'], 'launches', d['gpu_launches'])"
