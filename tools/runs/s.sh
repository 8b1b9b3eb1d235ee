#!/bin/bash
# full GPU suite file by file (a hang in one file cannot eat the whole call) + the default bench line
mkdir -p gpurun_out
T=gpurun_out/r2s_tests.log; : > $T
for f in tests/test_abi.py tests/test_gpu_gemm.py tests/test_gpu_metrics.py tests/test_gpu_panel.py tests/test_gpu_qr.py tests/test_gpu_solve.py tests/test_gpu_tsqr.py tests/test_loader.py tests/test_gpu_parity_large.py tests/test_gpu_mg.py; do
  echo "== $f" >> $T
  timeout -k 10 420 python -m pytest $f -m gpu -q --timeout 200 2>&1 | tail -15 >> $T
  echo "rc=${PIPESTATUS[0]}" >> $T
done
grep -E "^==|passed|failed|rc=" $T
timeout -k 10 400 python bench.py > gpurun_out/r2s_bench_c4.json 2> gpurun_out/r2s_bench_c4.err
echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2s_bench_c4.json; tail -3 gpurun_out/r2s_bench_c4.err
