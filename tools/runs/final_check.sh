#!/bin/bash
mkdir -p gpurun_out
W=tools/runs/watch.sh
rm -f gpurun_out/observed.jsonl
$W 200 gpurun_out/r2fin_test_gpu_qr.log python -m pytest tests/test_gpu_qr.py tests/test_gpu_panel.py -m gpu -q --timeout 0
tail -n 5 gpurun_out/r2fin_test_gpu_qr.log | cut -c1-400
timeout -k 10 200 python bench.py --no-cpu-baseline > gpurun_out/r2fin_bench_c4.json 2> gpurun_out/r2fin_bench_c4.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2fin_bench_c4.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline'])
PY
MPQR_NO_CHAIN=1 MPQR_OVERLAP=0 timeout -k 10 150 ncu --clock-control none --metrics gpu__time_duration.sum -c 7000 --csv --log-file gpurun_out/r2fin_launches_c4_serial.csv \
    python tools/quick_time.py 32768,32768,128,fp16,1024 > gpurun_out/r2fin_ncu_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r2fin_launches_c4_serial.csv)"; grep -v '^"' gpurun_out/r2fin_launches_c4_serial.csv | tail -n 4 | cut -c1-200
find gpurun_out -size +12M -print -delete; du -sm gpurun_out
