#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> <gpus> <script>   -- retries while the pod answers busy (exit 3)
T=$1; G=$2; S=$3
for i in $(seq 1 40); do
  if [ "$G" -gt 1 ]; then /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $S"; else /usr/local/graft/bin/gpurun --timeout $T -- "bash $S"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
