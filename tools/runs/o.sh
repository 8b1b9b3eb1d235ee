#!/bin/bash
mkdir -p gpurun_out
Q="timeout -k 5 60 python tools/quick_time.py 32768,32768,128,fp16"
MPQR_TRACE=1 $Q > gpurun_out/r2o_qt_default.log 2>&1; head -12 gpurun_out/r2o_qt_default.log
CHECK=0 MPQR_NO_GS_MERGE=1 $Q > gpurun_out/r2o_qt_nomerge.log 2>&1; head -1 gpurun_out/r2o_qt_nomerge.log
CHECK=0 MPQR_NO_REST_GATE=1 $Q > gpurun_out/r2o_qt_nogate.log 2>&1; head -1 gpurun_out/r2o_qt_nogate.log
CHECK=0 MPQR_REST_KEEP=16 $Q > gpurun_out/r2o_qt_keep16.log 2>&1; head -1 gpurun_out/r2o_qt_keep16.log
CHECK=0 MPQR_REST_KEEP=0 $Q > gpurun_out/r2o_qt_keep0.log 2>&1; head -1 gpurun_out/r2o_qt_keep0.log
CHECK=0 MPQR_NO_GS_MERGE=1 MPQR_NO_REST_GATE=1 MPQR_REST_KEEP=0 $Q > gpurun_out/r2o_qt_old.log 2>&1; head -1 gpurun_out/r2o_qt_old.log
timeout -k 10 100 python tools/timeline.py 32768,32768,128,fp16 20.0 20.9 > gpurun_out/r2o_timeline_c4.log 2>&1
sed -n 3,70p gpurun_out/r2o_timeline_c4.log | cut -c1-100
timeout -k 10 400 python -m pytest tests/test_gpu_qr.py tests/test_gpu_panel.py -x -q --timeout 100 > gpurun_out/r2o_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2o_tests.log
tail -5 gpurun_out/r2o_tests.log
