#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_qr.py -x -q --timeout 120 -k "lookahead or mixed_driver or larger or streamed or plan_cache" > gpurun_out/r2n_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2n_tests.log
tail -5 gpurun_out/r2n_tests.log
Q="timeout -k 10 100 python tools/quick_time.py 32768,32768,128,fp16"
MPQR_TRACE=1 $Q > gpurun_out/r2n_qt_default.log 2>&1; head -12 gpurun_out/r2n_qt_default.log
CHECK=0 MPQR_NO_GS_MERGE=1 $Q > gpurun_out/r2n_qt_nomerge.log 2>&1; head -1 gpurun_out/r2n_qt_nomerge.log
CHECK=0 MPQR_NO_REST_GATE=1 $Q > gpurun_out/r2n_qt_nogate.log 2>&1; head -1 gpurun_out/r2n_qt_nogate.log
CHECK=0 MPQR_REST_KEEP=16 $Q > gpurun_out/r2n_qt_keep16.log 2>&1; head -1 gpurun_out/r2n_qt_keep16.log
CHECK=0 MPQR_REST_KEEP=0 $Q > gpurun_out/r2n_qt_keep0.log 2>&1; head -1 gpurun_out/r2n_qt_keep0.log
CHECK=0 MPQR_NO_GS_MERGE=1 MPQR_NO_REST_GATE=1 MPQR_REST_KEEP=0 $Q > gpurun_out/r2n_qt_old.log 2>&1; head -1 gpurun_out/r2n_qt_old.log
timeout -k 10 200 python tools/timeline.py 32768,32768,128,fp16 20.0 20.9 > gpurun_out/r2n_timeline_c4.log 2>&1
sed -n 3,70p gpurun_out/r2n_timeline_c4.log | cut -c1-100
timeout -k 10 300 python tools/quick_time.py 2048,2048,32,fp16 4096,16384,64,fp16 8192,8192,128,fp16 16384,16384,128,fp16 > gpurun_out/r2n_qt_small.log 2>&1
cat gpurun_out/r2n_qt_small.log
