#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <script> [args]  — retries while the pod answers "transient / busy"
T=$1; shift
for i in $(seq 1 20); do
  out=$(gpurun --timeout $T -- bash "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy\|rc=3"; then
    echo "[retry $i] pod busy"; sleep 120; continue
  fi
  echo "$out" | tail -60
  exit 0
done
echo "gave up"
