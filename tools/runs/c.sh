#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/observed.jsonl
python tools/chain_probe.py 32768,128 24576,128 16384,128 8192,128 2048,128 > gpurun_out/r2c_chain_probe.log 2>&1
cat gpurun_out/r2c_chain_probe.log
timeout -k 10 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_qr.py tests/test_gpu_panel.py -x -q --timeout 120 > gpurun_out/r2c_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c_tests.log
tail -3 gpurun_out/r2c_tests.log
MPQR_TRACE=1 timeout -k 10 200 python tools/quick_time.py 32768,32768,128,fp16 > gpurun_out/r2c_qt_model.log 2>&1
for sms in 32 48 64; do
  MPQR_PANEL_SMS=$sms MPQR_TRACE=1 timeout -k 10 200 python tools/quick_time.py 32768,32768,128,fp16 > gpurun_out/r2c_qt_sms$sms.log 2>&1
done
CHECK=0 timeout -k 10 200 python tools/quick_time.py 32768,32768,128,fp16,2048 32768,32768,128,fp16,1536 32768,32768,128,fp16,512 > gpurun_out/r2c_qt_nb.log 2>&1
head -2 gpurun_out/r2c_qt_model.log gpurun_out/r2c_qt_sms*.log gpurun_out/r2c_qt_nb.log
timeout -k 10 900 python -m pytest tests/test_gpu_parity_large.py -q --timeout 300 > gpurun_out/r2c_parity_large.log 2>&1
echo "parity rc=$?" >> gpurun_out/r2c_parity_large.log
tail -30 gpurun_out/r2c_parity_large.log
