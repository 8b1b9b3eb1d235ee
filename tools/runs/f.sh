#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/observed.jsonl
timeout -k 10 400 python -m pytest tests/test_gpu_qr.py -x -q --timeout 200 -k "streamed or plan_cache or shim" > gpurun_out/r2f_host_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2f_host_tests.log
tail -15 gpurun_out/r2f_host_tests.log
MPQR_HOST_TRACE=1 timeout -k 10 300 python tools/e2e_time.py > gpurun_out/r2f_e2e.log 2>&1
cat gpurun_out/r2f_e2e.log
MPQR_NO_STREAM_IN=1 MPQR_HOST_TRACE=1 timeout -k 10 300 python tools/e2e_time.py > gpurun_out/r2f_e2e_plain.log 2>&1
tail -6 gpurun_out/r2f_e2e_plain.log
timeout -k 10 200 python tools/gemm_time.py > gpurun_out/r2f_gemm_time.log 2>&1
cat gpurun_out/r2f_gemm_time.log
MPQR_TRACE=1 timeout -k 10 200 python tools/quick_time.py 32768,32768,128,fp16 > gpurun_out/r2f_qt_model.log 2>&1
head -12 gpurun_out/r2f_qt_model.log
timeout -k 10 300 python tools/quick_time.py 2048,2048,32,fp16 4096,16384,64,fp16 8192,8192,128,fp16 16384,16384,128,fp16 > gpurun_out/r2f_qt_small.log 2>&1
cat gpurun_out/r2f_qt_small.log
timeout -k 10 200 python bench.py --workload c5 --steps 3 --warmup 3 > gpurun_out/r2f_bench_c5.log 2>&1
tail -c 600 gpurun_out/r2f_bench_c5.log
