#!/bin/bash
# executed warp instructions and DRAM bytes of the pair kernel, one launch per workload at FULL size
# (ncu, three metrics only): feeds profiles/k2_inst_per_pair.json and profiles/k2_dram_traffic.json
mkdir -p gpurun_out
for wl in target config3 config2 config5 config4 config1; do
timeout 900 ncu --metrics smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:pairs_tiled -c 3 --csv --log-file gpurun_out/r02w_counts_$wl.csv python bench.py --workload $wl --steps 1 --warmup 3 --quick > gpurun_out/r02w_counts_$wl.log 2>&1; echo $wl ncu_exit=$?
done
