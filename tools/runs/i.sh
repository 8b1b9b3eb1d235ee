#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python tools/chain_check.py 100 > gpurun_out/r2i_chain_check.log 2>&1
cat gpurun_out/r2i_chain_check.log
if grep -q "TIMEOUT\|FAILED" gpurun_out/r2i_chain_check.log; then echo "chain check failed: stopping"; exit 0; fi
timeout -k 10 500 python -m pytest tests/test_gpu_panel.py tests/test_gpu_qr.py -x -q --timeout 200 > gpurun_out/r2i_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2i_tests.log
tail -6 gpurun_out/r2i_tests.log
MPQR_TRACE=1 MPQR_HOST_TRACE=1 timeout -k 10 300 python tools/e2e_time.py > gpurun_out/r2i_e2e.log 2>&1
grep -v "^ *[0-9]" gpurun_out/r2i_e2e.log | tail -9
tail -37 gpurun_out/r2i_e2e.log | head -34
MPQR_NO_STREAM_IN=1 MPQR_HOST_TRACE=1 timeout -k 10 300 python tools/e2e_time.py > gpurun_out/r2i_e2e_plain.log 2>&1
grep -v "^ *[0-9]" gpurun_out/r2i_e2e_plain.log | tail -4
MPQR_PANEL_SMS=64 MPQR_TRACE=1 timeout -k 10 200 python tools/quick_time.py 32768,32768,128,fp16 > gpurun_out/r2i_qt64.log 2>&1
head -20 gpurun_out/r2i_qt64.log
MPQR_TRACE=1 timeout -k 10 200 python tools/quick_time.py 32768,32768,128,fp16 > gpurun_out/r2i_qt.log 2>&1
head -2 gpurun_out/r2i_qt.log
python tools/chain_probe.py 32768,128 16384,128 > gpurun_out/r2i_chain_probe.log 2>&1
cat gpurun_out/r2i_chain_probe.log
