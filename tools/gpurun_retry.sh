#!/bin/bash
# usage: tools/gpurun_retry.sh <gpus> <timeout> <command...>: retries while the pod answers busy (exit code 3 / "transient")
G=$1; T=$2; shift 2
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy\|no box"; then echo "[retry $i] busy"; sleep 150; continue; fi
  echo "$out"; exit 0
done
echo "gave up"; exit 3
