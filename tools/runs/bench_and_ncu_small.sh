#!/bin/bash
mkdir -p gpurun_out
W=tools/runs/watch.sh
timeout -k 10 400 python bench.py > gpurun_out/r2z_bench_c4.json 2> gpurun_out/r2z_bench_c4.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/r2z_bench_c4.json
timeout -k 10 200 python bench.py --impl reference > gpurun_out/r2z_bench_reference_arm.json 2> gpurun_out/r2z_bench_reference_arm.err
for w in c2 c3 c5; do
  timeout -k 10 150 python bench.py --workload $w --no-cpu-baseline > gpurun_out/r2z_bench_$w.json 2> gpurun_out/r2z_bench_$w.err
done
rm -f gpurun_out/observed.jsonl
$W 150 gpurun_out/r2z_tests_refshapes.log python -m pytest tests/test_gpu_qr.py -m gpu -q --timeout 0 -k "fp32_driver or mixed_driver or larger or lookahead_driver"
tail -n 3 gpurun_out/r2z_tests_refshapes.log; cp gpurun_out/observed.jsonl gpurun_out/r2z_observed_qr.jsonl
NCU="ncu --clock-control none"
timeout -k 10 200 $NCU --set full --import-source on -k regex:panel_chain_kernel -c 1 -f -o gpurun_out/r2z_chain python tools/chain_probe.py 32768,32 > gpurun_out/r2z_ncu_chain.log 2>&1
echo "chain rc=$?"
MPQR_NO_CHAIN=1 timeout -k 10 200 $NCU --set full -k regex:"inpanel_[su]4_kernel|panel_block_kernel" -c 6 -f -o /tmp/r2z_panel python tools/chain_probe.py 32768,128 > gpurun_out/r2z_ncu_panel.log 2>&1
echo "panel rc=$?"
timeout -k 10 300 $NCU --set full -k regex:tc_gemm2_kernel -c 7 -f -o /tmp/r2z_gemm python tools/gemm_time.py > gpurun_out/r2z_ncu_gemm.log 2>&1
echo "gemm rc=$?"
for r in /tmp/r2z_panel /tmp/r2z_gemm gpurun_out/r2z_chain; do
  [ -f $r.ncu-rep ] && ncu -i $r.ncu-rep --page raw --csv > gpurun_out/$(basename $r)_raw.csv 2>/dev/null
done
[ -f gpurun_out/r2z_chain.ncu-rep ] && ncu -i gpurun_out/r2z_chain.ncu-rep --page source --csv > gpurun_out/r2z_chain_source.csv 2>/dev/null
MPQR_NO_CHAIN=1 timeout -k 10 240 $NCU --metrics gpu__time_duration.sum -c 3000 --csv --log-file gpurun_out/r2z_launches_c4_stream_ordered.csv \
    python tools/quick_time.py 32768,32768,128,fp16,1024 > gpurun_out/r2z_ncu_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r2z_launches_c4_stream_ordered.csv)"; tail -n 5 gpurun_out/r2z_ncu_launches.log | cut -c1-300; tail -n 3 gpurun_out/r2z_launches_c4_stream_ordered.csv | cut -c1-300
find gpurun_out -size +12M -print -delete
du -sm gpurun_out
