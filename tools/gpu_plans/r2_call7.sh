#!/bin/bash
# round 2, call 7 (2 GPUs): everything multi-GPU once more on the current build — numerics worker (LL default, in-place
# trees), soak (10^4 graph-replayed mixed ops), ZeRO-1 parity, workflow with the coordinator process and the algorithm plan,
# the ViT / MoE workloads, bench variants, in-situ timelines at N=1 and N=2.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
ADAPCC_TIMEOUT_MS=15000 timeout 500 $TR --master-port 29601 tests/gpu_collectives_worker.py --quick --sweep --out gpurun_out/c7_worker.json > gpurun_out/c7_worker.log 2>&1
grep -E "FAIL|failures|checks per rank|Error|error" gpurun_out/c7_worker.log | head -12; grep -E "\[sweep\]" gpurun_out/c7_worker.log | cut -c1-150,380-520 | head -10
timeout 300 $TR --master-port 29611 tests/gpu_soak_worker.py --ops 10000 > gpurun_out/c7_soak.log 2>&1; grep -E "soak|Error|Traceback" gpurun_out/c7_soak.log | tail -6
ADAPCC_TIMEOUT_MS=15000 timeout 400 $TR --master-port 29602 tests/gpu_zero1_worker.py > gpurun_out/c7_zero1.log 2>&1; grep -E "zero1|Error|Traceback" gpurun_out/c7_zero1.log | tail -6
ADAPCC_TIMEOUT_MS=15000 timeout 300 $TR --master-port 29604 tests/gpu_workflow_worker.py > gpurun_out/c7_workflow.log 2>&1; grep -E "workflow\]|relay_steps|plan|Error|Traceback" gpurun_out/c7_workflow.log | tail -8
ADAPCC_TIMEOUT_MS=20000 timeout 300 $TR --master-port 29801 examples/train_vit.py --entry_point 7 --profile_freq 6 --steps 14 --batch 64 > gpurun_out/c7_vit.log 2>&1; grep -E "step (1|5|6|7|13) |reconstruct|Traceback|Error" gpurun_out/c7_vit.log | head -10
ADAPCC_TIMEOUT_MS=20000 timeout 300 $TR --master-port 29802 examples/train_moe.py --steps 12 --stragglers 1 --straggle_ms 100 > gpurun_out/c7_moe_relay.log 2>&1; tail -5 gpurun_out/c7_moe_relay.log | cut -c1-200
ADAPCC_TIMEOUT_MS=20000 timeout 300 $TR --master-port 29803 examples/train_moe.py --steps 8 --expert_parallel > gpurun_out/c7_moe_ep.log 2>&1; tail -4 gpurun_out/c7_moe_ep.log | cut -c1-200
b() { n=$1; shift; env "$@" timeout 300 $TR --master-port 29603 bench.py --gpus 2 --steps 20 --warmup 5 $EXTRA > gpurun_out/c7_bench2_$n.json 2> gpurun_out/c7_bench2_$n.err; echo "$n: $(tail -1 gpurun_out/c7_bench2_$n.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d.get("vs_baseline"), d.get("allreduce_check"), d.get("baseline_arm",{}).get("ms_per_step"))' 2>&1 | tail -1)"; tail -1 gpurun_out/c7_bench2_$n.err | cut -c1-200; }
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/c7_bench1.json 2> gpurun_out/c7_bench1.err; tail -1 gpurun_out/c7_bench1.json | cut -c1-230
EXTRA="" b default X=1
EXTRA="--zero1 --no_nccl_arm" b zero1 X=1
EXTRA="--engine ddp --no_nccl_arm" b ddp_hook X=1
EXTRA="--engine ddp --relay_control --no_nccl_arm" b ddp_hook_relay X=1
EXTRA="--engine ddp --impl nccl" b ddp_nccl X=1
timeout 200 python tools/torch_profile_ddp.py --out gpurun_out/c7_timeline_1.md > gpurun_out/c7_timeline1.log 2>&1; head -3 gpurun_out/c7_timeline_1.md | cut -c1-300
timeout 200 $TR --master-port 29605 tools/torch_profile_ddp.py --out gpurun_out/c7_timeline_2.md > gpurun_out/c7_timeline2.log 2>&1; head -3 gpurun_out/c7_timeline_2.md | cut -c1-300; tail -2 gpurun_out/c7_timeline2.log
timeout 200 $TR --master-port 29606 tools/torch_profile_ddp.py --zero1 --out gpurun_out/c7_timeline_2_zero1.md > gpurun_out/c7_timeline_z.log 2>&1; head -3 gpurun_out/c7_timeline_2_zero1.md | cut -c1-300
