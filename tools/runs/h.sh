#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_qr.py -x -q --timeout 200 -k "streamed or plan_cache" > gpurun_out/r2h_host_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2h_host_tests.log
tail -8 gpurun_out/r2h_host_tests.log
MPQR_TRACE=1 MPQR_HOST_TRACE=1 timeout -k 10 300 python tools/e2e_time.py > gpurun_out/r2h_e2e.log 2>&1
grep -v "^ *[0-9]" gpurun_out/r2h_e2e.log | tail -14
tail -37 gpurun_out/r2h_e2e.log | head -34
MPQR_NO_STREAM_IN=1 MPQR_HOST_TRACE=1 timeout -k 10 300 python tools/e2e_time.py > gpurun_out/r2h_e2e_plain.log 2>&1
grep -v "^ *[0-9]" gpurun_out/r2h_e2e_plain.log | tail -5
MPQR_PANEL_SMS=64 timeout -k 10 200 python tools/quick_time.py 32768,32768,128,fp16 > gpurun_out/r2h_qt.log 2>&1
head -2 gpurun_out/r2h_qt.log
