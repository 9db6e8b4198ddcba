#!/bin/bash
# Pipelined one-shot call: parity tests, then e2e with and without it on the three large workloads.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipelined or staged" > gpurun_out/r02q_pytest.log 2>&1
tail -3 gpurun_out/r02q_pytest.log
line() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; g=d.get('e2e_pageable') or {}
        print('$1', 'value %.4g ms %.2f | e2e %.4g ms %.2f | pageable %.4g ms %.2f' % (d['value'], d['ms_per_step'], e['value'], e['ms_per_step'], g.get('value',0), g.get('ms_per_step',0)))
"; }
for w in target config5 config3; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --steps 4 --warmup 3 2>gpurun_out/r02q_$w.err | tee gpurun_out/r02q_$w.json | line "$w pipe  "
  ICIKT_NO_PIPELINE=1 timeout 600 python bench.py --workload $w --no-cpu-baseline --steps 4 --warmup 3 2>/dev/null | tee gpurun_out/r02q_${w}_plain.json | line "$w plain "
done
