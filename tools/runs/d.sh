#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/observed.jsonl
timeout -k 10 300 python -m pytest tests/test_gpu_gemm.py -x -q --timeout 100 > gpurun_out/r2d_gemm_tests.log 2>&1
echo "gemm tests rc=$?" >> gpurun_out/r2d_gemm_tests.log
tail -5 gpurun_out/r2d_gemm_tests.log
timeout -k 10 200 python tools/gemm_time.py > gpurun_out/r2d_gemm_time.log 2>&1
timeout -k 10 200 python tools/gemm_time.py 16384 16384 1024 >> gpurun_out/r2d_gemm_time.log 2>&1
cat gpurun_out/r2d_gemm_time.log
if grep -q "failed\|error\|Error" gpurun_out/r2d_gemm_tests.log; then export MPQR_GEMM_1CTA=1; echo "2-CTA GEMM failed: continuing with the one-CTA kernels"; fi
timeout -k 10 600 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/r2d_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2d_tests.log
tail -8 gpurun_out/r2d_tests.log
MPQR_PANEL_SMS=64 MPQR_TRACE=1 timeout -k 10 200 python tools/quick_time.py 32768,32768,128,fp16 > gpurun_out/r2d_qt_sms64.log 2>&1
MPQR_PANEL_SMS=48 MPQR_TRACE=1 timeout -k 10 200 python tools/quick_time.py 32768,32768,128,fp16 > gpurun_out/r2d_qt_sms48.log 2>&1
MPQR_TRACE=1 timeout -k 10 200 python tools/quick_time.py 32768,32768,128,fp16 > gpurun_out/r2d_qt_model.log 2>&1
MPQR_OVERLAP=0 PROFILE=1 timeout -k 10 200 python tools/quick_time.py 32768,32768,128,fp16 > gpurun_out/r2d_qt_serial.log 2>&1
head -3 gpurun_out/r2d_qt_*.log
