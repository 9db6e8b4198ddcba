#!/bin/bash
# Sweep of the tie-group thresholds of K1 (ICIKT_LARGE_TIE: size from which a tie group may be
# "large", ICIKT_DIRECT_BUDGET: sum of size^2 / n up to which such groups are still compared
# directly) on the count-data workloads.  K2 time in ms.
out=gpurun_out/sweep_ties.txt
: > $out
run() { # name T B bench-args...
  name=$1; T=$2; B=$3; shift 3
  r=$(ICIKT_LARGE_TIE=$T ICIKT_DIRECT_BUDGET=$B timeout 300 python bench.py "$@" --steps 3 --warmup 3 --quick 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('k2_ms=%.4g k1_ms=%.3g' % (r['k2_ms'], r['k1_ms']))
")
  echo "$name T=$T B=$B $r" | tee -a $out
}
for tb in "128 24" "256 24" "256 48" "512 48" "192 32" "128 48"; do
  set -- $tb
  run yeast $1 $2 --workload config1
  run counts5000x300 $1 $2 --workload config4 --rows 5000 --cols 300
  run counts20000x300 $1 $2 --workload config4 --rows 20000 --cols 300
  run counts60000x60 $1 $2 --workload config4 --cols 60
done
