#!/bin/bash
mkdir -p gpurun_out
echo "== threads" > gpurun_out/r2r_tsqr.log
MPQR_HOST_TRACE=1 timeout -k 10 100 python tools/tsqr_time.py >> gpurun_out/r2r_tsqr.log 2>&1
echo "== one thread" >> gpurun_out/r2r_tsqr.log
MPQR_TSQR_ONE_THREAD=1 MPQR_HOST_TRACE=1 timeout -k 10 100 python tools/tsqr_time.py >> gpurun_out/r2r_tsqr.log 2>&1
echo "== threads lanes 8" >> gpurun_out/r2r_tsqr.log
MPQR_TSQR_LANES=8 timeout -k 10 100 python tools/tsqr_time.py >> gpurun_out/r2r_tsqr.log 2>&1
echo "== threads lanes 8 hmax 16384" >> gpurun_out/r2r_tsqr.log
MPQR_TSQR_LANES=8 MPQR_TSQR_HMAX=16384 timeout -k 10 100 python tools/tsqr_time.py >> gpurun_out/r2r_tsqr.log 2>&1
grep -v "^tsqr 8192" gpurun_out/r2r_tsqr.log | cut -c1-200
timeout -k 10 200 python tools/quick_time.py 2048,2048,32,fp16,512 2048,2048,32,fp16,256 4096,16384,64,fp16,512 4096,16384,64,fp16,256 8192,8192,128,fp16,512 > gpurun_out/r2r_qt_nb.log 2>&1
cat gpurun_out/r2r_qt_nb.log
timeout -k 10 1100 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2r_tests_all.log 2>&1
echo "rc=$?" >> gpurun_out/r2r_tests_all.log
tail -6 gpurun_out/r2r_tests_all.log
