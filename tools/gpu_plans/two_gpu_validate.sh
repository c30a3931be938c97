#!/bin/bash
# 2 GPUs: first runs of the LL all-reduce and of the sharded optimizer, then their effect on the bench.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
ADAPCC_LL=1 ADAPCC_EXPERIMENTAL=1 ADAPCC_TIMEOUT_MS=15000 timeout 400 $TR --master-port 29601 tests/gpu_collectives_worker.py --quick --sweep --out gpurun_out/p2_worker_ll.json > gpurun_out/p2_worker_ll.log 2>&1
grep -E "FAIL|failures|checks per rank" gpurun_out/p2_worker_ll.log | head -20; grep "\[sweep\]" gpurun_out/p2_worker_ll.log | head -8 | cut -c1-400
ADAPCC_TIMEOUT_MS=15000 timeout 300 $TR --master-port 29604 tests/gpu_workflow_worker.py > gpurun_out/p2_workflow.log 2>&1; grep -E "workflow\]|relay_steps" gpurun_out/p2_workflow.log | tail -5
ADAPCC_TIMEOUT_MS=15000 timeout 300 $TR --master-port 29602 tests/gpu_zero1_worker.py > gpurun_out/p2_zero1.log 2>&1; grep "zero1" gpurun_out/p2_zero1.log | tail -6
for v in "default:" "zero1:--zero1"; do n=${v%%:*}; f=${v#*:}; timeout 200 $TR --master-port 29603 bench.py --gpus 2 --steps 20 --warmup 5 $f > gpurun_out/p2_bench_$n.json 2> gpurun_out/p2_bench_$n.err; echo "$n: $(tail -1 gpurun_out/p2_bench_$n.json | cut -c1-330)"; done
