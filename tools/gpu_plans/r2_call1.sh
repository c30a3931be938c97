#!/bin/bash
# round 2, call 1 (1 GPU): full GPU test suite, first runs of tcgen05 variants 1/2, GEMM timing vs cuBLAS,
# both bench arms (ours + the reference's stock GPT-2 path), in-situ kernel table.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/c1_pytest.log 2>&1; tail -3 gpurun_out/c1_pytest.log
VARIANTS=0
for v in 1 2; do
  ADAPCC_TCGEN05_TEST_VARIANTS=$v timeout 150 python -m pytest tests/test_gpu_tcgen05.py -q -x > gpurun_out/c1_tcgen05_v$v.log 2>&1
  tail -3 gpurun_out/c1_tcgen05_v$v.log
  grep -q " passed" gpurun_out/c1_tcgen05_v$v.log && ! grep -q "failed\|error" gpurun_out/c1_tcgen05_v$v.log && VARIANTS=$VARIANTS,$v
done
echo "tcgen05 variants that pass: $VARIANTS"
timeout 150 python -m adapcc_b200.bench.gemm_bench --variants $VARIANTS --json gpurun_out/c1_gemm_bench.json > gpurun_out/c1_gemm_bench.log 2>&1; cat gpurun_out/c1_gemm_bench.log
timeout 150 python -m adapcc_b200.bench.gemm_bench --m 8192 --n 768 --k 3072 --variants $VARIANTS --json gpurun_out/c1_gemm_bench_proj.json > gpurun_out/c1_gemm_bench_proj.log 2>&1; cat gpurun_out/c1_gemm_bench_proj.log
run() { n=$1; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/c1_bench_$n.json 2> gpurun_out/c1_bench_$n.err; echo "$n: $(tail -1 gpurun_out/c1_bench_$n.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["clocks"])' 2>&1 | tail -1)"; }
run default X=1
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/c1_bench_reference.json 2> gpurun_out/c1_bench_reference.err; tail -1 gpurun_out/c1_bench_reference.json | cut -c1-400; tail -3 gpurun_out/c1_bench_reference.err
run tc_mlp_v0 ADAPCC_TCGEN05_MLP=1
case "$VARIANTS" in *1*) run tc_mlp_v1 ADAPCC_TCGEN05_MLP=1 ADAPCC_TCGEN05_VARIANT=1;; esac
case "$VARIANTS" in *2*) run tc_mlp_v2 ADAPCC_TCGEN05_MLP=1 ADAPCC_TCGEN05_VARIANT=2;; esac
run tc_mlp_fwd_bwd ADAPCC_TCGEN05_MLP=2
timeout 150 python tools/torch_profile_step.py --out gpurun_out/c1_torch_profile.md > gpurun_out/c1_torch_profile.log 2>&1; head -40 gpurun_out/c1_torch_profile.md
timeout 200 python bench.py --impl reference --ref_precision fp32 --steps 10 --warmup 3 > gpurun_out/c1_bench_reference_fp32.json 2> gpurun_out/c1_bench_reference_fp32.err; tail -1 gpurun_out/c1_bench_reference_fp32.json | cut -c1-200
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
