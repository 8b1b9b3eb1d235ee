#!/bin/bash
mkdir -p gpurun_out
W=tools/runs/watch.sh
$W 60 gpurun_out/r2v_qt_c4.log python tools/quick_time.py 32768,32768,128,fp16,1024
grep TFLOP gpurun_out/r2v_qt_c4.log | cut -c1-200
$W 90 gpurun_out/r2v_timeline_c4.log python tools/timeline.py 32768,32768,128,fp16 20.0 20.2
sed -n 3,16p gpurun_out/r2v_timeline_c4.log | cut -c1-120
for sched in "1,3,4,8" "1,3,4,4,4,8,4,2,2" "1,3,4,8,8,4,2,2" "1,1,2,4,4,4,8,4,2,2"; do
  echo "== chunks $sched" >> gpurun_out/r2v_e2e.log
  MPQR_H2D_CHUNKS=$sched MPQR_HOST_TRACE=1 timeout -k 5 80 python tools/e2e_time.py >> gpurun_out/r2v_e2e.log 2>&1
done
echo "== chunks default, ARR_SMS=80" >> gpurun_out/r2v_e2e.log
MPQR_ARR_SMS=80 MPQR_HOST_TRACE=1 timeout -k 5 80 python tools/e2e_time.py >> gpurun_out/r2v_e2e.log 2>&1
grep -E "^==|^call|chunk arr|backward" gpurun_out/r2v_e2e.log | cut -c1-330
$W 200 gpurun_out/r2v_tests.log python -m pytest tests/test_gpu_panel.py tests/test_gpu_qr.py -q --timeout 0 -k "panel or lookahead or streamed or stream_ordered or larger"
tail -n 4 gpurun_out/r2v_tests.log | cut -c1-300
$W 60 gpurun_out/r2v_timeline_c5.log python tools/timeline.py tsqr,1048576,256 0 0
sed -n 3,60p gpurun_out/r2v_timeline_c5.log | cut -c1-120
