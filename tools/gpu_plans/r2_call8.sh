#!/bin/bash
# round 2, call 8 (2 GPUs, short): ZeRO-1 diagnostics (per-slice mismatch against an NCCL reference), 6-stage variant 3
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
ADAPCC_TIMEOUT_MS=15000 timeout 400 $TR --master-port 29602 tests/gpu_zero1_worker.py > gpurun_out/c8_zero1.log 2>&1; grep -E "zero1" gpurun_out/c8_zero1.log | cut -c1-330 | head -40
timeout 300 python -m pytest tests/test_gpu_tcgen05_pp.py -q -x > gpurun_out/c8_pp_tests.log 2>&1; tail -3 gpurun_out/c8_pp_tests.log
timeout 200 python -m adapcc_b200.bench.gemm_bench --variants 3 --json gpurun_out/c8_gemm_3072.json > gpurun_out/c8_gemm_3072.log 2>&1; tail -8 gpurun_out/c8_gemm_3072.log
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/c8_bench1.json 2> gpurun_out/c8_bench1.err; tail -1 gpurun_out/c8_bench1.json | cut -c1-230
ADAPCC_TCGEN05_MLP=0 timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/c8_bench1_cublas.json 2> gpurun_out/c8_bench1_cublas.err; tail -1 gpurun_out/c8_bench1_cublas.json | cut -c1-230
