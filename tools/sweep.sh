#!/bin/bash
# Tuning sweep of the pair kernel's launch shape (warps per CTA, register class) per workload.
out=gpurun_out/sweep.txt
: > $out
run() { # workload W R
  r=$(ICIKT_WARPS=$2 ICIKT_REGCLASS=$3 timeout 300 python bench.py --workload $1 --steps 3 --warmup 3 --quick 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('%.4g pairs/s k2_ms=%.4g k1_ms=%.3g frac=%.3f' % (d['value'], r['k2_ms'], r['k1_ms'], r['frac']))
")
  echo "$1 W=$2 R=$3 $r" | tee -a $out
}
for W in "" 2 3 4 5 6 7 10; do run config2 "$W" ""; done
run config2 4 0; run config2 4 1; run config2 4 2; run config2 3 0
for W in "" 1 2 3; do run config5 "$W" ""; done
run config5 1 0; run config5 1 1; run config5 1 2
for W in "" 9 12 16 20 27; do run target "$W" ""; done
run target 16 0; run target 16 1; run target 16 2
