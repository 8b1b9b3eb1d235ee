#!/bin/bash
mkdir -p gpurun_out
echo "== default" > gpurun_out/r2q_tsqr.log
timeout -k 10 100 python tools/tsqr_time.py >> gpurun_out/r2q_tsqr.log 2>&1
echo "== MPQR_INKERNEL=1" >> gpurun_out/r2q_tsqr.log
MPQR_INKERNEL=1 timeout -k 10 100 python tools/tsqr_time.py >> gpurun_out/r2q_tsqr.log 2>&1
echo "== MPQR_INKERNEL=1 lanes 8" >> gpurun_out/r2q_tsqr.log
MPQR_INKERNEL=1 MPQR_TSQR_LANES=8 timeout -k 10 100 python tools/tsqr_time.py >> gpurun_out/r2q_tsqr.log 2>&1
echo "== MPQR_GATE_KERNEL=1" >> gpurun_out/r2q_tsqr.log
MPQR_GATE_KERNEL=1 timeout -k 10 100 python tools/tsqr_time.py >> gpurun_out/r2q_tsqr.log 2>&1
cat gpurun_out/r2q_tsqr.log
MPQR_HOST_TRACE=1 timeout -k 10 120 python tools/e2e_time.py > gpurun_out/r2q_e2e_default.log 2>&1
grep -v "^ *[0-9]" gpurun_out/r2q_e2e_default.log | tail -4
MPQR_INKERNEL=48 MPQR_HOST_TRACE=1 timeout -k 10 120 python tools/e2e_time.py > gpurun_out/r2q_e2e_ink.log 2>&1
grep -v "^ *[0-9]" gpurun_out/r2q_e2e_ink.log | tail -4
MPQR_INKERNEL=32 MPQR_HOST_TRACE=1 timeout -k 10 120 python tools/e2e_time.py > gpurun_out/r2q_e2e_ink32.log 2>&1
grep -v "^ *[0-9]" gpurun_out/r2q_e2e_ink32.log | tail -4
MPQR_INKERNEL=48 timeout -k 10 500 python -m pytest tests/test_gpu_qr.py tests/test_gpu_panel.py tests/test_gpu_tsqr.py -x -q --timeout 100 > gpurun_out/r2q_tests_ink.log 2>&1
echo "rc=$?" >> gpurun_out/r2q_tests_ink.log
tail -5 gpurun_out/r2q_tests_ink.log
