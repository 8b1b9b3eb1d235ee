#!/bin/bash
# first validation of the persistent panel chain: panel + driver parity, then timing with traces
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.log 2>&1
timeout -k 10 400 python -m pytest tests/test_gpu_panel.py -x -q > gpurun_out/r2a_panel.log 2>&1
echo "panel rc=$?" >> gpurun_out/r2a_panel.log
timeout -k 10 600 python -m pytest tests/test_gpu_qr.py -x -q > gpurun_out/r2a_qr.log 2>&1
echo "qr rc=$?" >> gpurun_out/r2a_qr.log
MPQR_TRACE=1 timeout -k 10 300 python tools/quick_time.py 32768,32768,128,fp16 > gpurun_out/r2a_qt_chain.log 2>&1
echo "rc=$?" >> gpurun_out/r2a_qt_chain.log
MPQR_NO_CHAIN=1 MPQR_TRACE=1 timeout -k 10 300 python tools/quick_time.py 32768,32768,128,fp16 > gpurun_out/r2a_qt_nochain.log 2>&1
echo "rc=$?" >> gpurun_out/r2a_qt_nochain.log
MPQR_GATE_KERNEL=1 MPQR_TRACE=1 timeout -k 10 300 python tools/quick_time.py 32768,32768,128,fp16 > gpurun_out/r2a_qt_gate.log 2>&1
echo "rc=$?" >> gpurun_out/r2a_qt_gate.log
MPQR_OVERLAP=0 PROFILE=1 timeout -k 10 300 python tools/quick_time.py 32768,32768,128,fp16 16384,16384,128,fp16 8192,8192,128,fp16 > gpurun_out/r2a_qt_serial.log 2>&1
echo "rc=$?" >> gpurun_out/r2a_qt_serial.log
timeout -k 10 300 python tools/quick_time.py 2048,2048,32,fp16 4096,16384,64,fp16 8192,8192,128,fp16 16384,16384,128,fp16 > gpurun_out/r2a_qt_small.log 2>&1
tail -3 gpurun_out/r2a_panel.log gpurun_out/r2a_qr.log gpurun_out/r2a_qt_chain.log
