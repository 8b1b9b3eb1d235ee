#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 200 python tools/h2d_probe.py > gpurun_out/r2g_h2d_probe.log 2>&1
cat gpurun_out/r2g_h2d_probe.log
MPQR_TRACE=1 MPQR_HOST_TRACE=1 timeout -k 10 300 python tools/e2e_time.py > gpurun_out/r2g_e2e.log 2>&1
tail -60 gpurun_out/r2g_e2e.log
