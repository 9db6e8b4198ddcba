#!/bin/bash
# Tuning sweep of the pair kernel's launch shape (warps per CTA, register class) per workload.
out=gpurun_out/sweep.txt
: > $out
for w in config2 config5 target; do
  for W in 2 4 8 16 32; do
    for R in 0 1 2; do
      if [ "$w" = "target" ] && [ $W -lt 16 ]; then continue; fi
      if [ "$w" = "config5" ] && [ $W -gt 8 ]; then continue; fi
      if [ $R -eq 2 ] && [ $W -gt 8 ]; then continue; fi
      if [ $R -eq 1 ] && [ $W -gt 16 ]; then continue; fi
      r=$(ICIKT_WARPS=$W ICIKT_REGCLASS=$R timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --quick 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('%.4g pairs/s k2_ms=%.4g frac=%.3f' % (d['value'], r['k2_ms'], r['frac']))
")
      echo "$w W=$W R=$R $r" | tee -a $out
    done
  done
done
