#!/bin/bash
# round 2, call 2 (2 GPUs): new 1-GPU kernels (fused embedding, CE scale), then everything multi-GPU that never ran:
# LL all-reduce, reduce-scatter/all-gather, tail-pack fix, ZeRO-1, adopt_native; bench N=2 variants; in-situ timeline.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x > gpurun_out/c2_pytest_ops.log 2>&1; tail -4 gpurun_out/c2_pytest_ops.log
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/c2_bench1.json 2> gpurun_out/c2_bench1.err; tail -1 gpurun_out/c2_bench1.json | cut -c1-260; tail -2 gpurun_out/c2_bench1.err
ADAPCC_FUSED_EMBED=0 timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/c2_bench1_noembed.json 2> gpurun_out/c2_bench1_noembed.err; tail -1 gpurun_out/c2_bench1_noembed.json | cut -c1-200
ADAPCC_TIMEOUT_MS=15000 timeout 500 $TR --master-port 29601 tests/gpu_collectives_worker.py --quick --sweep --out gpurun_out/c2_worker.json > gpurun_out/c2_worker.log 2>&1
grep -E "FAIL|failures|checks per rank|Error|error" gpurun_out/c2_worker.log | head -20; grep -E "\[sweep\]|\[prims\]" gpurun_out/c2_worker.log | cut -c1-420 | head -24
ADAPCC_TIMEOUT_MS=15000 timeout 300 $TR --master-port 29604 tests/gpu_workflow_worker.py > gpurun_out/c2_workflow.log 2>&1; grep -E "workflow\]|relay_steps|Error|Traceback" gpurun_out/c2_workflow.log | tail -6
ADAPCC_TIMEOUT_MS=15000 timeout 300 $TR --master-port 29602 tests/gpu_zero1_worker.py > gpurun_out/c2_zero1.log 2>&1; grep -E "zero1|Error|Traceback" gpurun_out/c2_zero1.log | tail -8
b() { n=$1; shift; env "$@" timeout 300 $TR --master-port 29603 bench.py --gpus 2 --steps 20 --warmup 5 $EXTRA > gpurun_out/c2_bench2_$n.json 2> gpurun_out/c2_bench2_$n.err; echo "$n: $(tail -1 gpurun_out/c2_bench2_$n.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d.get("vs_baseline"), d.get("allreduce_check"), d.get("baseline_arm",{}).get("ms_per_step"))' 2>&1 | tail -1)"; }
EXTRA="" b default X=1
EXTRA="--no_nccl_arm" b blocks8 ADAPCC_MAX_BLOCKS=8
EXTRA="--no_nccl_arm" b blocks16 ADAPCC_MAX_BLOCKS=16
EXTRA="--no_nccl_arm" b blocks32 ADAPCC_MAX_BLOCKS=32
EXTRA="--zero1 --no_nccl_arm" b zero1 X=1
EXTRA="--impl reference" b reference X=1
timeout 200 $TR --master-port 29605 tools/torch_profile_ddp.py --out gpurun_out/c2_timeline_2.md > gpurun_out/c2_timeline.log 2>&1; head -30 gpurun_out/c2_timeline_2.md; tail -3 gpurun_out/c2_timeline.log
timeout 200 $TR --master-port 29606 tools/torch_profile_ddp.py --zero1 --out gpurun_out/c2_timeline_2_zero1.md > gpurun_out/c2_timeline_z.log 2>&1; head -12 gpurun_out/c2_timeline_2_zero1.md; tail -3 gpurun_out/c2_timeline_z.log
