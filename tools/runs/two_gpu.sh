#!/bin/bash
mkdir -p gpurun_out
W=tools/runs/watch.sh
$W 200 gpurun_out/r2mg_tests.log python -m pytest tests/test_gpu_mg.py -m gpu -q --timeout 0
tail -n 4 gpurun_out/r2mg_tests.log | cut -c1-300
timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 > gpurun_out/r2mg_bench_c4_2gpu.json 2> gpurun_out/r2mg_bench_c4_2gpu.err
echo "c4 rc=$?"; grep '^{' gpurun_out/r2mg_bench_c4_2gpu.json | cut -c1-400
timeout -k 10 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --workload c5 > gpurun_out/r2mg_bench_c5_2gpu.json 2> gpurun_out/r2mg_bench_c5_2gpu.err
echo "c5 rc=$?"; grep '^{' gpurun_out/r2mg_bench_c5_2gpu.json | cut -c1-400
du -sm gpurun_out
