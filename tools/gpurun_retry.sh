#!/usr/bin/env bash
# usage: tools/gpurun_retry.sh <timeout_s> <gpus> '<command>' <logfile>   — retries while the pod answers busy/transient
to=$1; gpus=$2; cmd=$3; log=$4
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  if [ "$gpus" = "1" ]; then gpurun --timeout "$to" -- "$cmd" > "$log" 2>&1; else gpurun --gpus "$gpus" --timeout "$to" -- "$cmd" > "$log" 2>&1; fi
  if grep -q "status=transient\|rc=3\|retry in a few minutes\|no box\|busy" "$log" && ! grep -q "status=ok\|status=fail" "$log"; then sleep 120; continue; fi
  break
done
