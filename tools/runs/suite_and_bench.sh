#!/bin/bash
# full GPU suite file by file (a hang in one file cannot eat the whole call), smoke, and the bench lines of every workload
mkdir -p gpurun_out
W=tools/runs/watch.sh
T=gpurun_out/r2y_tests.log; : > $T
for f in test_abi test_gpu_gemm test_gpu_metrics test_gpu_panel test_gpu_qr test_gpu_solve test_gpu_tsqr test_loader test_gpu_parity_large test_gpu_mg; do
  echo "== $f" >> $T
  rm -f gpurun_out/observed.jsonl
  $W 400 gpurun_out/r2y_$f.log python -m pytest tests/$f.py -m gpu -q --timeout 0
  tail -n 6 gpurun_out/r2y_$f.log >> $T
  [ -f gpurun_out/observed.jsonl ] && cp gpurun_out/observed.jsonl gpurun_out/r2y_observed_$f.jsonl
done
grep -E "^==|passed|failed|rc=" $T
timeout -k 10 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2y_smoke.log 2>&1; tail -n 3 gpurun_out/r2y_smoke.log
timeout -k 10 400 python bench.py > gpurun_out/r2y_bench_c4.json 2> gpurun_out/r2y_bench_c4.err
echo "bench rc=$?"; cut -c1-900 gpurun_out/r2y_bench_c4.json
timeout -k 10 200 python bench.py --impl reference > gpurun_out/r2y_bench_reference_arm.json 2> gpurun_out/r2y_bench_reference_arm.err; cut -c1-300 gpurun_out/r2y_bench_reference_arm.json
for w in c2 c3 c5; do
  timeout -k 10 150 python bench.py --workload $w --no-cpu-baseline > gpurun_out/r2y_bench_$w.json 2> gpurun_out/r2y_bench_$w.err
  cut -c1-260 gpurun_out/r2y_bench_$w.json; echo
done
