#!/bin/bash
# one batch: (1) default (stream memops) chain flow on the streamed host path again, (2) fast SGEMM + TSQR, (3) timelines and sweeps
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_qr.py -x -q --timeout 120 -k "streamed or plan_cache" > gpurun_out/r2l_host_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2l_host_tests.log
tail -4 gpurun_out/r2l_host_tests.log
MPQR_HOST_TRACE=1 timeout -k 10 120 python tools/e2e_time.py > gpurun_out/r2l_e2e.log 2>&1
grep -v "^ *[0-9]" gpurun_out/r2l_e2e.log | tail -6
bash tools/runs/k.sh
bash tools/runs/j.sh
MPQR_INKERNEL=1 timeout -k 10 100 python tools/quick_time.py 32768,32768,128,fp16 > gpurun_out/r2l_qt_inkernel.log 2>&1
head -3 gpurun_out/r2l_qt_inkernel.log
