#!/bin/bash
# 8 GPUs: the headline numbers. Expensive (charged x8): run only what the smaller plans have already validated.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
LL=${ADAPCC_LL:-0}
ADAPCC_LL=$LL ADAPCC_TIMEOUT_MS=20000 timeout 900 $TR --master-port 29701 tests/gpu_collectives_worker.py --sweep --out gpurun_out/p8_sweep.json > gpurun_out/p8_sweep.log 2>&1
grep -E "FAIL|failures|checks per rank" gpurun_out/p8_sweep.log | head; grep "\[sweep\]" gpurun_out/p8_sweep.log | cut -c1-330 | head -24
for v in "default:" "zero1:--zero1" "ddp_hook:--engine ddp" "ddp_nccl:--engine ddp --impl nccl"; do n=${v%%:*}; f=${v#*:}; [ "$n" = "zero1" ] && [ "${ZERO1:-0}" != "1" ] && continue; timeout 300 $TR --master-port 29702 bench.py --gpus 8 --steps 20 --warmup 5 $f > gpurun_out/p8_bench_$n.json 2> gpurun_out/p8_bench_$n.err; echo "$n: $(tail -1 gpurun_out/p8_bench_$n.json | cut -c1-330)"; done
