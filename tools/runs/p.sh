#!/bin/bash
mkdir -p gpurun_out
Q="timeout -k 5 60 python tools/quick_time.py 32768,32768,128,fp16"
run() { name=$1; shift; env CHECK=0 REPS=4 "$@" $Q 2>&1 | head -1 | sed "s/^/$name: /" >> gpurun_out/r2p_ab.log; }
for rep in 1 2; do
  run "merged nogate keep0 " MPQR_REST_KEEP=0
  run "merged nogate keep32" MPQR_REST_KEEP=32
  run "merged gate   keep32" MPQR_REST_KEEP=32
  run "merged gate   keep48" MPQR_REST_KEEP=48
  run "plain  nogate keep0 " MPQR_NO_GS_MERGE=1 MPQR_REST_KEEP=0
  run "plain  nogate keep32" MPQR_NO_GS_MERGE=1 MPQR_REST_KEEP=32
done
cat gpurun_out/r2p_ab.log
timeout -k 10 100 python tools/timeline.py 32768,32768,128,fp16 20.0 20.3 > gpurun_out/r2p_timeline_c4.log 2>&1
sed -n 3,22p gpurun_out/r2p_timeline_c4.log | cut -c1-100
timeout -k 10 300 python -m pytest tests/test_gpu_qr.py -x -q --timeout 100 -k "lookahead or larger" > gpurun_out/r2p_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2p_tests.log
tail -3 gpurun_out/r2p_tests.log
