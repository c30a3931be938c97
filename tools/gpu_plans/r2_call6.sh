#!/bin/bash
# round 2, call 6 (1 GPU): re-validation after the epilogue / embedding-scatter changes, timing, ncu, heap diagnostic
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tcgen05_pp.py tests/test_gpu_ops.py -q -x > gpurun_out/c6_tests.log 2>&1; tail -4 gpurun_out/c6_tests.log
for shape in "8192 3072 768" "8192 2304 768"; do set -- $shape; timeout 200 python -m adapcc_b200.bench.gemm_bench --m $1 --n $2 --k $3 --variants 3 --json gpurun_out/c6_gemm_$2.json > gpurun_out/c6_gemm_$2.log 2>&1; tail -8 gpurun_out/c6_gemm_$2.log; done
run() { n=$1; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/c6_bench_$n.json 2> gpurun_out/c6_bench_$n.err; echo "$n: $(tail -1 gpurun_out/c6_bench_$n.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["clocks"]["reasons"])' 2>&1 | tail -1)"; tail -2 gpurun_out/c6_bench_$n.err | cut -c1-300; }
run default_tc3 X=1
run cublas_mlp ADAPCC_TCGEN05_MLP=0
run fuse_add_ln ADAPCC_FUSE_ADD_LN=1
run default_tc3_again X=1
timeout 150 python tools/torch_profile_step.py --out gpurun_out/c6_torch_profile.md > gpurun_out/c6_torch_profile.log 2>&1; grep -E "adapcc::|kernel time" gpurun_out/c6_torch_profile.md | head -20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_pair_persistent' -c 3 \
  -o gpurun_out/c6_ncu_gemm -f python tools/ncu_targets.py gemm 3 > gpurun_out/c6_ncu_gemm.log 2>&1; tail -2 gpurun_out/c6_ncu_gemm.log
