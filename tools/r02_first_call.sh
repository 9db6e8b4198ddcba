#!/bin/bash
# round 2, GPU call 1: parity suite (incl. the full-size configs), the default bench (target, N=1),
# the strong-scaled bench on 2 GPUs, and the ncu launch list of the default command.
export ICIKT_REQUIRE_GPU=1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_gpus.txt
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_pytest_gpu.log 2>&1; echo pytest_exit=$? >> gpurun_out/r02_pytest_gpu.log
tail -15 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02_bench_target_n1.json 2> gpurun_out/r02_bench_target_n1.err; echo bench_exit=$?
tail -c 3000 gpurun_out/r02_bench_target_n1.json
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02_bench_target_n2.json 2> gpurun_out/r02_bench_target_n2.err; echo bench2_exit=$?
tail -c 2500 gpurun_out/r02_bench_target_n2.json
tail -5 gpurun_out/r02_bench_target_n2.err
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_target.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/r02_ncu_launches.log 2>&1; echo ncu_exit=$?
