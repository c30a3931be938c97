#!/bin/bash
# N GPUs (default 4): the two BASELINE.json workload configs that only the GPT-2 bench has numbers for so far.
#   config 4: ViT DDP with reconstruct_topology on the profiling path (profile_freq shortened so it triggers)
#   config 5: MoE DDP with relay control, two ranks forced late (they become relays)
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
ADAPCC_TIMEOUT_MS=20000 timeout 400 $TR --master-port 29801 examples/train_vit.py --entry_point 7 --profile_freq 15 --steps 40 \
  > gpurun_out/pw_vit_$N.log 2>&1; grep -E "step (1|14|15|16|39) |reconstruct|Traceback|Error" gpurun_out/pw_vit_$N.log | head -12
S=$((N-1)); [ $N -ge 8 ] && S="$((N-2)),$((N-1))"
ADAPCC_TIMEOUT_MS=20000 timeout 400 $TR --master-port 29802 examples/train_moe.py --steps 30 --stragglers $S --straggle_ms 100 \
  > gpurun_out/pw_moe_relay_$N.log 2>&1; tail -6 gpurun_out/pw_moe_relay_$N.log
ADAPCC_TIMEOUT_MS=20000 timeout 400 $TR --master-port 29803 examples/train_moe.py --steps 30 --expert_parallel \
  > gpurun_out/pw_moe_ep_$N.log 2>&1; tail -4 gpurun_out/pw_moe_ep_$N.log
