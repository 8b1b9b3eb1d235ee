#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 240 python tools/timeline.py 32768,32768,128,fp16 20.0 21.4 > gpurun_out/r2j_timeline_c4.log 2>&1
head -40 gpurun_out/r2j_timeline_c4.log
for nb in 1280 1536 2048; do
  CHECK=0 REPS=3 timeout -k 10 200 python tools/quick_time.py 32768,32768,128,fp16,$nb >> gpurun_out/r2j_qt_nb.log 2>&1
done
cat gpurun_out/r2j_qt_nb.log
timeout -k 10 200 python tools/chain_probe.py 2048,128 2048,128,1,8 2048,128,2,4 4096,128 4096,128,2,8 8192,128,4,8 > gpurun_out/r2j_chain_probe.log 2>&1
grep -v "block [1-6]" gpurun_out/r2j_chain_probe.log
MPQR_MIN_R=128 timeout -k 10 200 python tools/quick_time.py 2048,2048,32,fp16 4096,16384,64,fp16 > gpurun_out/r2j_qt_minr.log 2>&1
cat gpurun_out/r2j_qt_minr.log
timeout -k 10 240 python tools/timeline.py 2048,2048,32,fp16 0.5 1.2 > gpurun_out/r2j_timeline_c2.log 2>&1
head -30 gpurun_out/r2j_timeline_c2.log
timeout -k 10 240 python tools/timeline.py 4096,16384,64,fp16 2.0 2.6 > gpurun_out/r2j_timeline_c3.log 2>&1
head -30 gpurun_out/r2j_timeline_c3.log
