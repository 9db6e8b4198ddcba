#!/bin/bash
export ICIKT_REQUIRE_GPU=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 900 -k "config4 or long_vectors or random_parity or uneven or yeast or small_matrices or collisions" > gpurun_out/r02o_pytest.log 2>&1; tail -3 gpurun_out/r02o_pytest.log
for wl in config4 config1; do
timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --quick 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$wl', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'k1', round(r['k1_ms'],3), 'k2', round(r['k2_ms'],3), 'frac', round(r['frac'],3))"
done
timeout 300 python bench.py --workload config4 --rows 20000 --cols 300 --steps 5 --warmup 3 --quick 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('counts n=20000 x300', round(d['value']), 'k2', round(r['k2_ms'],3), 'frac', round(r['frac'],3))"
