#!/bin/bash
mkdir -p gpurun_out
W=tools/runs/watch.sh
$W 330 gpurun_out/r2u_parity_large.log python -m pytest tests/test_gpu_parity_large.py -q --timeout 0
tail -n 12 gpurun_out/r2u_parity_large.log | cut -c1-300
cp gpurun_out/observed.jsonl gpurun_out/r2u_observed_parity_large.jsonl 2>/dev/null
$W 120 gpurun_out/r2u_tsqr_tests.log python -m pytest tests/test_gpu_tsqr.py tests/test_gpu_qr.py -q --timeout 0 -k "tsqr or stream_ordered"
tail -n 5 gpurun_out/r2u_tsqr_tests.log | cut -c1-300
timeout -k 5 60 python tools/tsqr_time.py > gpurun_out/r2u_tsqr.log 2>&1; cut -c1-200 gpurun_out/r2u_tsqr.log
timeout -k 5 120 python bench.py --workload c5 > gpurun_out/r2u_bench_c5.json 2> gpurun_out/r2u_bench_c5.err; cut -c1-600 gpurun_out/r2u_bench_c5.json
$W 45 gpurun_out/r2u_hazard_ordered.log python tools/alloc_hazard.py ordered
cat gpurun_out/r2u_hazard_ordered.log
$W 45 gpurun_out/r2u_hazard_chain.log python tools/alloc_hazard.py
cat gpurun_out/r2u_hazard_chain.log
grep -A4 "Kernel Parent" gpurun_out/r2u_*.gdb 2>/dev/null | cut -c1-220
