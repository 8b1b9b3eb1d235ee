#!/bin/bash
mkdir -p gpurun_out
python tools/chain_check.py 100 > gpurun_out/r2b_chain_check.log 2>&1
cat gpurun_out/r2b_chain_check.log
if grep -q "TIMEOUT\|FAILED" gpurun_out/r2b_chain_check.log; then echo "chain check failed: stopping"; exit 0; fi
timeout -k 10 500 python -m pytest tests/test_gpu_panel.py tests/test_gpu_qr.py -x -q --timeout 120 > gpurun_out/r2b_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2b_tests.log
tail -5 gpurun_out/r2b_tests.log
MPQR_TRACE=1 timeout -k 10 200 python tools/quick_time.py 32768,32768,128,fp16 > gpurun_out/r2b_qt_chain.log 2>&1
echo "rc=$?" >> gpurun_out/r2b_qt_chain.log
MPQR_NO_CHAIN=1 MPQR_TRACE=1 timeout -k 10 200 python tools/quick_time.py 32768,32768,128,fp16 > gpurun_out/r2b_qt_nochain.log 2>&1
MPQR_GATE_KERNEL=1 timeout -k 10 200 python tools/quick_time.py 32768,32768,128,fp16 > gpurun_out/r2b_qt_gate.log 2>&1
MPQR_OVERLAP=0 PROFILE=1 timeout -k 10 200 python tools/quick_time.py 32768,32768,128,fp16 16384,16384,128,fp16 > gpurun_out/r2b_qt_serial.log 2>&1
timeout -k 10 200 python tools/quick_time.py 2048,2048,32,fp16 4096,16384,64,fp16 8192,8192,128,fp16 16384,16384,128,fp16 > gpurun_out/r2b_qt_small.log 2>&1
tail -4 gpurun_out/r2b_qt_chain.log gpurun_out/r2b_qt_nochain.log gpurun_out/r2b_qt_gate.log
