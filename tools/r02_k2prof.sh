#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pairs_tiled -c 1 -o gpurun_out/r02z_k2_target300 -f python bench.py --steps 1 --warmup 3 --quick --cols 300 > gpurun_out/r02z_ncu_k2.log 2>&1; echo ncu_k2_exit=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z_launches_target.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/r02z_ncu_launches.log 2>&1; echo ncu_exit=$?
