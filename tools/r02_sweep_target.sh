#!/bin/bash
# launch-shape sweep of the pair kernel at the target shape (300 samples) after the round-2 changes
out=gpurun_out/r02_sweep_target.txt
: > $out
run() { # W R
  r=$(ICIKT_WARPS=$1 ICIKT_REGCLASS=$2 timeout 300 python bench.py --workload target --cols 400 --steps 3 --warmup 3 --quick 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('%.4g pairs/s k2_ms=%.4g k1_ms=%.3g frac=%.3f' % (d['value'], r['k2_ms'], r['k1_ms'], r['frac']))
")
  echo "target x400 W=$1 R=$2 $r" | tee -a $out
}
run "" ""; run 16 0; run 16 1; run 16 2; run 27 2; run 27 0; run 13 0; run 13 1; run 9 0; run 9 1; run 32 2
