#!/bin/bash
# round 2, the 8-GPU call (charged x8: only what the 1- and 2-GPU calls validated). Numerics + sweeps (all-reduce 1 KB-1 GB
# incl. LL / in-place trees; reduce / broadcast / all-to-all) next to NCCL, soak, bench (ours + in-process NCCL arm, ZeRO-1,
# the reference arm), in-situ timeline, BASELINE configs 4 (ViT + reconstruct_topology) and 5 (MoE, ranks 6,7 late -> relays),
# expert-parallel MoE, straggler bench (relay negotiation latency with the coordinator process).
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
T0=$(date +%s); lap() { echo "== $1 done at +$(( $(date +%s) - T0 )) s"; }
ADAPCC_TIMEOUT_MS=20000 timeout 420 $TR --master-port 29701 tests/gpu_collectives_worker.py --quick --sweep --sweep_max_log2 30 --sweep_step 2 --out gpurun_out/c9_worker_$N.json > gpurun_out/c9_worker_$N.log 2>&1
grep -E "FAIL|failures|checks per rank|Error" gpurun_out/c9_worker_$N.log | head -8; grep -E "\[sweep\]" gpurun_out/c9_worker_$N.log | cut -c1-100 | head -12; lap worker
timeout 200 $TR --master-port 29711 tests/gpu_soak_worker.py --ops 10000 > gpurun_out/c9_soak_$N.log 2>&1; grep -E "soak|Error|Traceback" gpurun_out/c9_soak_$N.log | tail -4; lap soak
timeout 150 $TR --master-port 29721 tests/gpu_engine_parity_worker.py > gpurun_out/c9_parity_$N.log 2>&1; grep -E "parity|Error|Traceback" gpurun_out/c9_parity_$N.log | head -6
ADAPCC_TIMEOUT_MS=15000 timeout 240 $TR --master-port 29722 tests/gpu_zero1_worker.py > gpurun_out/c9_zero1_$N.log 2>&1; grep -E "zero1\] (rank 0|failures)" gpurun_out/c9_zero1_$N.log | cut -c1-260 | head -6; lap parity
b() { n=$1; shift; env "$@" timeout 240 $TR --master-port 29703 bench.py --gpus $N --steps 20 --warmup 5 $EXTRA > gpurun_out/c9_bench${N}_$n.json 2> gpurun_out/c9_bench${N}_$n.err; echo "$n: $(tail -1 gpurun_out/c9_bench${N}_$n.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d.get("vs_baseline"), d.get("allreduce_check"), d.get("replicas_identical"), d.get("baseline_arm",{}).get("ms_per_step"))' 2>&1 | tail -1)"; tail -1 gpurun_out/c9_bench${N}_$n.err | cut -c1-200; }
EXTRA="" b default X=1
EXTRA="--no_zero1 --no_nccl_arm" b replicated X=1
EXTRA="--impl reference" b reference X=1; lap bench
timeout 150 $TR --master-port 29705 tools/torch_profile_ddp.py --out gpurun_out/c9_timeline_$N.md > gpurun_out/c9_timeline_$N.log 2>&1; head -3 gpurun_out/c9_timeline_$N.md | cut -c1-300
timeout 150 $TR --master-port 29706 tools/torch_profile_ddp.py --no_zero1 --out gpurun_out/c9_timeline_${N}_replicated.md > gpurun_out/c9_timeline_${N}z.log 2>&1; head -3 gpurun_out/c9_timeline_${N}_replicated.md | cut -c1-300; lap timeline
ADAPCC_TIMEOUT_MS=20000 timeout 240 $TR --master-port 29801 examples/train_vit.py --entry_point 7 --profile_freq 6 --steps 14 --batch 128 > gpurun_out/c9_vit_$N.log 2>&1; grep -E "step (1|5|6|7|13) |reconstruct|Traceback|Error" gpurun_out/c9_vit_$N.log | head -10; lap vit
S=$((N-1)); [ $N -ge 8 ] && S="$((N-2)),$((N-1))"
ADAPCC_TIMEOUT_MS=20000 timeout 200 $TR --master-port 29802 examples/train_moe.py --steps 14 --stragglers $S --straggle_ms 100 --relay_mode forward --algo tree > gpurun_out/c9_moe_relay_$N.log 2>&1; grep -E "rank 0\] step (1|5|13)|relay_steps|Error|Traceback" gpurun_out/c9_moe_relay_$N.log | cut -c1-160 | tail -12
ADAPCC_TIMEOUT_MS=20000 timeout 200 $TR --master-port 29803 examples/train_moe.py --steps 8 --expert_parallel > gpurun_out/c9_moe_ep_$N.log 2>&1; tail -3 gpurun_out/c9_moe_ep_$N.log | cut -c1-200; lap moe
timeout 200 $TR --master-port 29804 -m adapcc_b200.bench.straggler_bench --stragglers $S --straggle_ms 100 --out gpurun_out/c9_straggler_$N.json > gpurun_out/c9_straggler_$N.log 2>&1; tail -4 gpurun_out/c9_straggler_$N.log | cut -c1-250; lap straggler
