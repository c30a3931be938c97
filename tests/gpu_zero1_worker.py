"""torchrun worker: the sharded-optimizer engine mode (reduce-scatter + AdamW-with-broadcast, csrc/zero.cu) against
plain data parallelism (all-reduce + replicated AdamW) on the same seeds, eager and under CUDA-graph capture.

    torchrun --nproc-per-node 2 tests/gpu_zero1_worker.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from adapcc_b200.models.gpt2 import GPT2Config, GPT2DoubleHeads, synthetic_batch  # noqa: E402
from adapcc_b200.parallel.engine import FlatDataParallel  # noqa: E402
from adapcc_b200.runtime.native import NativeComm  # noqa: E402
from adapcc_b200.runtime.rendezvous import unique_name  # noqa: E402


def run(comm, rank, world, dev, zero1, graph, steps=8):
    cfg = GPT2Config(vocab_size=1000, n_positions=64, n_embd=256, n_layer=2, n_head=4, lm_chunk_rows=128)
    torch.manual_seed(7)
    model = GPT2DoubleHeads(cfg).to(dev)
    comm.heap_reset()
    eng = FlatDataParallel(model, comm, world_size=world, rank=rank, lr=2e-3, max_norm=1.0, bucket_mb=0.5,
                           zero1=zero1)
    assert eng.zero1 == zero1
    batches = [synthetic_batch(2, 2, 64, cfg.vocab_size, device=dev, seed=100 * rank + i) for i in range(3)]
    losses = []
    if graph:
        eng.capture(batches[0], warmup=1)
        for i in range(steps):
            losses.append(float(eng.step_graph(batches[i % 3]).item()))
    else:
        for i in range(steps):
            losses.append(float(eng.step(batches[i % 3]).item()))
    comm.check()
    params = eng.flat_param.float().clone()
    eng.close()
    return losses, params


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    comm = NativeComm(unique_name("zero1"), rank, world, local, staging_bytes=16 << 20, heap_bytes=64 << 20)
    ok = True
    for graph in (False, True):
        base_l, base_p = run(comm, rank, world, dev, zero1=False, graph=graph)
        z_l, z_p = run(comm, rank, world, dev, zero1=True, graph=graph)
        # every rank must hold the same parameters after the broadcast
        ref = z_p.clone()
        dist.broadcast(ref, src=0)
        same = bool(torch.equal(ref, z_p))
        drift = float((z_p - base_p).abs().max())
        loss_gap = abs(z_l[-1] - base_l[-1]) / abs(base_l[-1])
        good = same and loss_gap < 0.03 and base_l[-1] < base_l[0] - 0.1 and z_l[-1] < z_l[0] - 0.1
        ok &= good
        if rank == 0 or not good:
            print(f"[zero1] rank {rank} graph={graph}: baseline {base_l[0]:.3f}->{base_l[-1]:.3f}, "
                  f"zero1 {z_l[0]:.3f}->{z_l[-1]:.3f}, replicas identical={same}, max param drift {drift:.3g} "
                  f"{'OK' if good else 'FAIL'}", flush=True)
    t = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(t)
    if rank == 0:
        print(f"[zero1] failures: {int(t.item())}", flush=True)
    dist.barrier()
    comm.close()
    dist.destroy_process_group()
    sys.exit(1 if t.item() else 0)


if __name__ == "__main__":
    main()
