#!/bin/bash
# round 2, call 11 (2 GPUs, short): sanity of the defaults that go into the 8-GPU call — ZeRO-1 default, embedding tables in
# their own bucket, deterministic grad norm, async relay descriptors
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29621 tests/gpu_engine_parity_worker.py > gpurun_out/c11_parity.log 2>&1; grep -E "parity|Error|Traceback" gpurun_out/c11_parity.log | head -6
ADAPCC_TIMEOUT_MS=15000 timeout 300 $TR --master-port 29602 tests/gpu_zero1_worker.py > gpurun_out/c11_zero1.log 2>&1; grep -E "zero1\] (rank 0|failures)" gpurun_out/c11_zero1.log | cut -c1-260 | head -6
b() { n=$1; shift; env "$@" timeout 300 $TR --master-port 29603 bench.py --gpus 2 --steps 20 --warmup 5 $EXTRA > gpurun_out/c11_bench2_$n.json 2> gpurun_out/c11_bench2_$n.err; echo "$n: $(tail -1 gpurun_out/c11_bench2_$n.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d.get("vs_baseline"), d.get("allreduce_check"), d.get("replicas_identical"), d["config"]["zero1"], d["config"]["buckets"], d.get("baseline_arm",{}).get("ms_per_step"))' 2>&1 | tail -1)"; tail -1 gpurun_out/c11_bench2_$n.err | cut -c1-200; }
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/c11_bench1.json 2> gpurun_out/c11_bench1.err; tail -1 gpurun_out/c11_bench1.json | cut -c1-230
EXTRA="" b default X=1
EXTRA="--no_zero1 --no_nccl_arm" b replicated X=1
timeout 200 $TR --master-port 29605 tools/torch_profile_ddp.py --out gpurun_out/c11_timeline_2.md > gpurun_out/c11_timeline2.log 2>&1; head -12 gpurun_out/c11_timeline_2.md | cut -c1-300
timeout 300 python -m pytest tests -q -m gpu -x > gpurun_out/c11_pytest.log 2>&1; tail -3 gpurun_out/c11_pytest.log
