#!/bin/bash
mkdir -p gpurun_out
W=tools/runs/watch.sh
$W 80 gpurun_out/r2t_c5_alone.log python -m pytest tests/test_gpu_parity_large.py -q -k "c5_full" --timeout 0
tail -3 gpurun_out/r2t_c5_alone.log
MPQR_NO_CHAIN=1 $W 80 gpurun_out/r2t_c5_nochain.log python -m pytest tests/test_gpu_parity_large.py -q -k "c5_full" --timeout 0
tail -3 gpurun_out/r2t_c5_nochain.log
$W 150 gpurun_out/r2t_c5_seq.log python -m pytest tests/test_gpu_parity_large.py -q -k "c3_full or c5_full or lookahead_tall" --timeout 0
tail -3 gpurun_out/r2t_c5_seq.log
MPQR_NO_CHAIN=1 timeout -k 5 60 python tools/tsqr_time.py > gpurun_out/r2t_tsqr_nochain.log 2>&1; cat gpurun_out/r2t_tsqr_nochain.log | cut -c1-200
MPQR_NO_CHAIN=1 MPQR_TSQR_LANES=8 timeout -k 5 60 python tools/tsqr_time.py > gpurun_out/r2t_tsqr_nochain8.log 2>&1; cat gpurun_out/r2t_tsqr_nochain8.log | cut -c1-200
head -c 3000 gpurun_out/*.gdb 2>/dev/null
