#!/bin/bash
mkdir -p gpurun_out
W=tools/runs/watch.sh
$W 60 gpurun_out/r2w_qt_c4.log python tools/quick_time.py 32768,32768,128,fp16,1024
grep TFLOP gpurun_out/r2w_qt_c4.log | cut -c1-200
MPQR_STREAM_POST=1 $W 60 gpurun_out/r2w_qt_c4_streampost.log python tools/quick_time.py 32768,32768,128,fp16,1024
grep TFLOP gpurun_out/r2w_qt_c4_streampost.log | cut -c1-200
MPQR_SU_MAXROWS=704 $W 60 gpurun_out/r2w_qt_c4_704.log python tools/quick_time.py 32768,32768,128,fp16,1024
grep TFLOP gpurun_out/r2w_qt_c4_704.log | cut -c1-200
$W 60 gpurun_out/r2w_qt_small.log python tools/quick_time.py 2048,2048,32,fp16,0 4096,16384,64,fp16,0 8192,8192,128,fp16,0 16384,16384,128,fp16,0
grep TFLOP gpurun_out/r2w_qt_small.log | cut -c1-200
$W 200 gpurun_out/r2w_tests.log python -m pytest tests/test_gpu_panel.py tests/test_gpu_qr.py -q --timeout 0 -k "panel or lookahead or streamed or stream_ordered or larger"
tail -n 4 gpurun_out/r2w_tests.log | cut -c1-300
for L in 12 16; do MPQR_TSQR_LANES=$L timeout -k 5 60 python tools/tsqr_time.py 2>&1 | grep -E "^tsqr 1048576|orth" | cut -c1-200 | sed "s/^/lanes $L: /" >> gpurun_out/r2w_tsqr_lanes.log; done
cat gpurun_out/r2w_tsqr_lanes.log
