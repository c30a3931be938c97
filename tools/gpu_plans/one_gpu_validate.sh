#!/bin/bash
# 1 GPU: validate what has only been compiled so far, then measure it.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu > gpurun_out/p1_pytest.log 2>&1; tail -3 gpurun_out/p1_pytest.log
# separate processes: a trapped experimental kernel must not poison the main suite's (or the other variant's) context
VARIANTS=0
for v in 1 2; do
  ADAPCC_EXPERIMENTAL=1 ADAPCC_TCGEN05_TEST_VARIANTS=$v timeout 120 python -m pytest tests/test_gpu_tcgen05.py -q -x > gpurun_out/p1_tcgen05_v$v.log 2>&1
  tail -3 gpurun_out/p1_tcgen05_v$v.log
  grep -q " passed" gpurun_out/p1_tcgen05_v$v.log && ! grep -q "failed\|error" gpurun_out/p1_tcgen05_v$v.log && VARIANTS=$VARIANTS,$v
done
echo "tcgen05 variants that pass: $VARIANTS"
timeout 120 python -m adapcc_b200.bench.gemm_bench --variants $VARIANTS --json gpurun_out/p1_gemm_bench.json > gpurun_out/p1_gemm_bench.log 2>&1; cat gpurun_out/p1_gemm_bench.log
timeout 120 python -m adapcc_b200.bench.gemm_bench --m 8192 --n 768 --k 3072 --variants $VARIANTS --json gpurun_out/p1_gemm_bench_proj.json > gpurun_out/p1_gemm_bench_proj.log 2>&1; cat gpurun_out/p1_gemm_bench_proj.log
run() { n=$1; shift; env "$@" timeout 150 python bench.py --steps 20 --warmup 5 > gpurun_out/p1_bench_$n.json 2> gpurun_out/p1_bench_$n.err; echo "$n: $(tail -1 gpurun_out/p1_bench_$n.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["clocks"]["reasons"])' 2>&1 | tail -1)"; }
run default X=1
run tc_mlp_v0 ADAPCC_TCGEN05_MLP=1
case "$VARIANTS" in *1*) run tc_mlp_v1 ADAPCC_TCGEN05_MLP=1 ADAPCC_TCGEN05_VARIANT=1;; esac
case "$VARIANTS" in *2*) run tc_mlp_v2 ADAPCC_TCGEN05_MLP=1 ADAPCC_TCGEN05_VARIANT=2;; esac
grep -q " passed" gpurun_out/p1_tcgen05_v1.log && run tc_mlp_fwd_bwd ADAPCC_TCGEN05_MLP=2
timeout 120 python tools/torch_profile_step.py --out gpurun_out/p1_torch_profile.md > gpurun_out/p1_torch_profile.log 2>&1; head -30 gpurun_out/p1_torch_profile.md
