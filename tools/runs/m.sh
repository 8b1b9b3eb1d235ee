#!/bin/bash
mkdir -p gpurun_out
Q="timeout -k 10 100 python tools/quick_time.py 32768,32768,128,fp16"
CHECK=0 MPQR_TRACE=1 $Q > gpurun_out/r2m_qt_default.log 2>&1; head -12 gpurun_out/r2m_qt_default.log
CHECK=0 MPQR_SU_MAXROWS=704 $Q > gpurun_out/r2m_qt_rows704.log 2>&1; head -1 gpurun_out/r2m_qt_rows704.log
CHECK=0 MPQR_INKERNEL=48 MPQR_SU_MAXROWS=704 $Q > gpurun_out/r2m_qt_ink48.log 2>&1; head -1 gpurun_out/r2m_qt_ink48.log
CHECK=0 MPQR_INKERNEL=48 $Q > gpurun_out/r2m_qt_ink48b.log 2>&1; head -1 gpurun_out/r2m_qt_ink48b.log
CHECK=0 MPQR_PANEL_SMS=80 $Q > gpurun_out/r2m_qt_sms80.log 2>&1; head -1 gpurun_out/r2m_qt_sms80.log
CHECK=0 MPQR_PANEL_SMS=80 MPQR_SU_MAXROWS=704 $Q > gpurun_out/r2m_qt_sms80_704.log 2>&1; head -1 gpurun_out/r2m_qt_sms80_704.log
timeout -k 10 200 python tools/timeline.py 32768,32768,128,fp16 20.0 21.0 > gpurun_out/r2m_timeline_c4.log 2>&1
sed -n 3,22p gpurun_out/r2m_timeline_c4.log
timeout -k 10 300 python tools/quick_time.py 2048,2048,32,fp16 4096,16384,64,fp16 8192,8192,128,fp16 16384,16384,128,fp16 > gpurun_out/r2m_qt_small.log 2>&1
cat gpurun_out/r2m_qt_small.log
timeout -k 10 600 python -m pytest tests/test_gpu_qr.py tests/test_gpu_panel.py -x -q --timeout 120 > gpurun_out/r2m_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2m_tests.log
tail -5 gpurun_out/r2m_tests.log
