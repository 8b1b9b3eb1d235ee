#!/bin/bash
# round-2 evidence: ncu captures (stream-ordered flow: under ncu kernels are serialised, the persistent panel kernel would
# wait for its side-stream work forever) + the chain kernel on its own (a 2-block panel has no side updates)
mkdir -p gpurun_out
NCU="ncu --clock-control none"
MPQR_NO_CHAIN=1 timeout -k 10 300 $NCU --metrics gpu__time_duration.sum -c 6600 --csv --log-file gpurun_out/r2x_launches_c4_stream_ordered.csv \
    python tools/quick_time.py 32768,32768,128,fp16,1024 > gpurun_out/r2x_ncu_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r2x_launches_c4_stream_ordered.csv)"
timeout -k 10 200 $NCU --set full --import-source on -k regex:panel_chain_kernel -c 1 -f -o gpurun_out/r2x_chain \
    python tools/chain_probe.py 32768,32 > gpurun_out/r2x_ncu_chain.log 2>&1
echo "chain rc=$?"; tail -n 3 gpurun_out/r2x_ncu_chain.log | cut -c1-200
MPQR_NO_CHAIN=1 timeout -k 10 200 $NCU --set full -k regex:"inpanel_[su]4_kernel|panel_finalize_kernel|tinv_kernel|panel_block_kernel" -c 30 -f -o gpurun_out/r2x_panel \
    python tools/chain_probe.py 32768,128 > gpurun_out/r2x_ncu_panel.log 2>&1
echo "panel rc=$?"
timeout -k 10 300 $NCU --set full -k regex:tc_gemm2_kernel -c 7 -f -o gpurun_out/r2x_gemm \
    python tools/gemm_time.py > gpurun_out/r2x_ncu_gemm.log 2>&1
echo "gemm rc=$?"; ls -la gpurun_out/*.ncu-rep
