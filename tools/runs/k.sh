#!/bin/bash
# FP32 fast SGEMM + restructured TSQR: correctness, then timing sweeps
mkdir -p gpurun_out
timeout -k 10 200 python tools/sgemm_time.py > gpurun_out/r2k_sgemm.log 2>&1
cat gpurun_out/r2k_sgemm.log
MPQR_SGEMM_OLD=1 timeout -k 10 200 python tools/sgemm_time.py > gpurun_out/r2k_sgemm_old.log 2>&1
head -8 gpurun_out/r2k_sgemm_old.log
timeout -k 10 600 python -m pytest tests/test_gpu_tsqr.py tests/test_gpu_solve.py tests/test_gpu_qr.py -x -q --timeout 300 -k "tsqr or solve or fp32" > gpurun_out/r2k_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2k_tests.log
tail -6 gpurun_out/r2k_tests.log
for lanes in 4 6 8; do for hm in 32768 16384; do
  echo "== lanes=$lanes hmax=$hm" >> gpurun_out/r2k_tsqr_time.log
  MPQR_TSQR_LANES=$lanes MPQR_TSQR_HMAX=$hm timeout -k 10 120 python tools/tsqr_time.py >> gpurun_out/r2k_tsqr_time.log 2>&1
done; done
cat gpurun_out/r2k_tsqr_time.log
