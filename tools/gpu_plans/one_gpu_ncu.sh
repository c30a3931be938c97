#!/bin/bash
# 1 GPU, ncu only (never a multi-rank command; numbers printed under ncu are never bench values).
mkdir -p gpurun_out
V=${1:-0,1}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_bias_act_tcgen05|nvjet' -c 12 \
  -o gpurun_out/ncu_gemm -f python tools/ncu_targets.py gemm $V > gpurun_out/ncu_gemm.log 2>&1; tail -3 gpurun_out/ncu_gemm.log
ncu -i gpurun_out/ncu_gemm.ncu-rep --page raw --csv > gpurun_out/ncu_gemm_raw.csv 2>/dev/null
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_gpt2.csv \
  python bench.py --engine eager --steps 2 --warmup 1 > gpurun_out/launches_bench.log 2>&1; tail -2 gpurun_out/launches_bench.log
