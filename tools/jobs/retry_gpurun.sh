#!/bin/bash
# usage: tools/jobs/retry_gpurun.sh <timeout-s> <gpus> <script-in-repo> ; retries while the pod answers "transient"/busy (rc 3)
T=$1; G=$2; S=$3
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "bash $S" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $S" 2>&1); fi
  RC=$?
  echo "$OUT" | tail -40
  if echo "$OUT" | grep -q "status=transient"; then echo "[retry $i] transient, sleeping"; sleep 90; continue; fi
  if [ $RC -eq 3 ]; then echo "[retry $i] rc=3, sleeping"; sleep 90; continue; fi
  exit $RC
done
exit 3
