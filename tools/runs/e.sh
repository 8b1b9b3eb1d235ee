#!/bin/bash
# 2 GPUs: multi-GPU parity tests + C4 / C5 bench lines at N = 2, then single-GPU checks of the new GEMM pipeline depth
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2e_smi.log
timeout -k 10 600 python -m pytest tests/test_gpu_mg.py -x -q --timeout 400 > gpurun_out/r2e_mg_tests.log 2>&1
echo "mg tests rc=$?" >> gpurun_out/r2e_mg_tests.log
tail -15 gpurun_out/r2e_mg_tests.log
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2e_bench_c4_2gpu.log 2> gpurun_out/r2e_bench_c4_2gpu.err
tail -c 1500 gpurun_out/r2e_bench_c4_2gpu.log; tail -5 gpurun_out/r2e_bench_c4_2gpu.err
timeout -k 10 200 python tools/gemm_time.py > gpurun_out/r2e_gemm_time.log 2>&1
cat gpurun_out/r2e_gemm_time.log
MPQR_TRACE=1 timeout -k 10 200 python tools/quick_time.py 32768,32768,128,fp16 > gpurun_out/r2e_qt_model.log 2>&1
head -40 gpurun_out/r2e_qt_model.log
