#!/bin/bash
# Round evidence: default bench line, reference arm, other workloads, ncu launch list and one
# --set full capture of the pair kernel per workload.  Usage: bash tools/evidence.sh <tag>
tag=${1:-r01}
set -x
python bench.py > gpurun_out/bench_${tag}_config2.json 2> gpurun_out/bench_${tag}_config2.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2> gpurun_out/bench_${tag}_reference.err
python bench.py --workload config1 --steps 10 --warmup 3 --cpu-budget 5 > gpurun_out/bench_${tag}_config1.json 2> gpurun_out/bench_${tag}_config1.err
for wl in config3 config4 config5 target; do
python bench.py --workload $wl --steps 3 --warmup 3 --cpu-budget 10 > gpurun_out/bench_${tag}_$wl.json 2> gpurun_out/bench_${tag}_$wl.err
done
python bench.py --kernel naive --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}_naive_config2.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}_config2.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/ncu_launch.log 2>&1
prof() { # name workload cols  (three pair-kernel launches per step, one per tier; all three of the timed step are captured, two of them exit at once)
ncu --set full --clock-control none --import-source on -k regex:pairs_tiled --launch-skip 9 --launch-count 3 -f -o gpurun_out/prof_${tag}_k2_$1 python bench.py --workload $2 --cols $3 --steps 1 --warmup 3 --quick > gpurun_out/ncu_$1.log 2>&1
}
prof config2 config2 100
prof target target 300
prof config1 config1 96
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
