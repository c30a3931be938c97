#!/bin/bash
# round 2, call 4 (1 GPU): first run of tcgen05 variant 3 (own process: a protocol bug traps the context), its timing
# against cuBLAS and variants 0-2, ncu of variants 1/3 and the cuBLAS kernels, GPT-2 bench with the fused MLP.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tcgen05_pp.py -q -x > gpurun_out/c4_pp_tests.log 2>&1; tail -5 gpurun_out/c4_pp_tests.log
PP_OK=0; grep -q " passed" gpurun_out/c4_pp_tests.log && ! grep -q "failed\|error" gpurun_out/c4_pp_tests.log && PP_OK=1
V=0,1,2; [ $PP_OK = 1 ] && V=0,1,2,3
timeout 200 python -m adapcc_b200.bench.gemm_bench --variants $V --json gpurun_out/c4_gemm_bench.json > gpurun_out/c4_gemm_bench.log 2>&1; cat gpurun_out/c4_gemm_bench.log | tail -16
timeout 200 python -m adapcc_b200.bench.gemm_bench --m 8192 --n 768 --k 3072 --variants $V --json gpurun_out/c4_gemm_bench_proj.json > gpurun_out/c4_gemm_bench_proj.log 2>&1; tail -10 gpurun_out/c4_gemm_bench_proj.log
timeout 200 python -m adapcc_b200.bench.gemm_bench --m 8192 --n 2304 --k 768 --variants 3 --json gpurun_out/c4_gemm_bench_qkv.json > gpurun_out/c4_gemm_bench_qkv.log 2>&1; tail -8 gpurun_out/c4_gemm_bench_qkv.log
run() { n=$1; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/c4_bench_$n.json 2> gpurun_out/c4_bench_$n.err; echo "$n: $(tail -1 gpurun_out/c4_bench_$n.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["clocks"]["reasons"])' 2>&1 | tail -1)"; tail -2 gpurun_out/c4_bench_$n.err | cut -c1-300; }
run default X=1
[ $PP_OK = 1 ] && run tc4_fwd ADAPCC_TCGEN05_MLP=1 ADAPCC_TCGEN05_VARIANT=3
[ $PP_OK = 1 ] && run tc4_fwd_bwd ADAPCC_TCGEN05_MLP=2 ADAPCC_TCGEN05_VARIANT=3
run fuse_add_ln ADAPCC_FUSE_ADD_LN=1
timeout 150 python tools/torch_profile_step.py --out gpurun_out/c4_torch_profile.md > gpurun_out/c4_torch_profile.log 2>&1; head -48 gpurun_out/c4_torch_profile.md | tail -44
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_pair_persistent' -c 6 \
  -o gpurun_out/c4_ncu_gemm -f python tools/ncu_targets.py gemm $([ $PP_OK = 1 ] && echo 1,3 || echo 1) > gpurun_out/c4_ncu_gemm.log 2>&1; tail -3 gpurun_out/c4_ncu_gemm.log
ncu -i gpurun_out/c4_ncu_gemm.ncu-rep --page raw --csv > gpurun_out/c4_ncu_gemm_raw.csv 2>/dev/null; wc -l gpurun_out/c4_ncu_gemm_raw.csv
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/c4_pytest.log 2>&1; tail -3 gpurun_out/c4_pytest.log
