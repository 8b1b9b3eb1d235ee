#!/bin/bash
# round-2 ncu evidence.  Under ncu kernels are serialised, so the persistent panel kernel (which waits for side-stream work) can
# only be captured on its own: a 2-block panel has no side updates.  Everything else is captured in the stream-ordered flow.
# Reports go to /tmp and come back as CSV (gpurun_out/ must stay under 64 MiB or NOTHING is copied back).
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout -k 10 200 $NCU --set full --import-source on -k regex:panel_chain_kernel -c 1 -f -o gpurun_out/ncu_chain python tools/chain_probe.py 32768,32 > gpurun_out/ncu_chain.log 2>&1
echo "chain rc=$?"
MPQR_NO_CHAIN=1 timeout -k 10 200 $NCU --set full -k regex:"inpanel_[su]4_kernel|panel_block_kernel" -c 6 -f -o /tmp/ncu_panel python tools/chain_probe.py 32768,128 > gpurun_out/ncu_panel.log 2>&1
echo "panel rc=$?"
timeout -k 10 300 $NCU --set full -k regex:tc_gemm2_kernel -c 7 -f -o /tmp/ncu_gemm python tools/gemm_time.py > gpurun_out/ncu_gemm.log 2>&1
echo "gemm rc=$?"
for r in /tmp/ncu_panel /tmp/ncu_gemm gpurun_out/ncu_chain; do
  [ -f $r.ncu-rep ] && ncu -i $r.ncu-rep --page raw --csv > gpurun_out/$(basename $r)_raw.csv 2>/dev/null
done
[ -f gpurun_out/ncu_chain.ncu-rep ] && ncu -i gpurun_out/ncu_chain.ncu-rep --page source --csv > gpurun_out/ncu_chain_source.csv 2>/dev/null
# launch list: serial stream-ordered flow (ncu cannot profile kernels launched inside a green-context partition)
MPQR_NO_CHAIN=1 MPQR_OVERLAP=0 timeout -k 10 400 $NCU --metrics gpu__time_duration.sum -c 7000 --csv --log-file gpurun_out/launches_c4_serial.csv \
    python tools/quick_time.py 32768,32768,128,fp16,1024 > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches_c4_serial.csv)"
find gpurun_out -size +12M -print -delete
du -sm gpurun_out
# summaries: python tools/ncu_summary.py gpurun_out/ncu_*_raw.csv
