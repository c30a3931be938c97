timeout 240 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_v3.log 2>&1; tail -12 gpurun_out/pytest_gpu_v3.log
run() { n=$1; shift; env "$@" timeout 120 python bench.py --steps 12 --warmup 3 $EXTRA > gpurun_out/bench_v3_$n.json 2> gpurun_out/bench_v3_$n.err; echo "$n: $(tail -1 gpurun_out/bench_v3_$n.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"])' 2>&1 | tail -1)"; }
EXTRA="" run default X=1
EXTRA="" run nodirect ADAPCC_DIRECT_GRADS=0
EXTRA="" run fuse ADAPCC_FUSE_ADD_LN=1
EXTRA="--lm_rows scored" run fuse_scored ADAPCC_FUSE_ADD_LN=1
ADAPCC_FUSE_ADD_LN=1 timeout 120 python tools/torch_profile_step.py --out gpurun_out/torch_profile_v3.md > gpurun_out/torch_profile_v3.log 2>&1; head -34 gpurun_out/torch_profile_v3.md
timeout 60 adapcc_b200/_C/check_p2p --mb 256 --iters 3 > gpurun_out/check_p2p_1gpu.log 2>&1; tail -8 gpurun_out/check_p2p_1gpu.log
