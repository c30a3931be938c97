"""torchrun worker: data-parallel gradient parity of the flat engine. One step with the native bucket all-reduce (sinks,
hooks, side stream) must leave in the flat gradient buffer exactly the average of the ranks' LOCAL gradients (computed
by a communication-free engine, averaged with NCCL) — per parameter, eager and under CUDA-graph capture. Catches buckets
launched before all of their gradients exist (round 2: every sink parameter was counted twice)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from adapcc_b200.models.gpt2 import GPT2Config, GPT2DoubleHeads, synthetic_batch  # noqa: E402
from adapcc_b200.parallel.engine import FlatDataParallel  # noqa: E402
from adapcc_b200.runtime.native import NativeComm  # noqa: E402
from adapcc_b200.runtime.rendezvous import unique_name  # noqa: E402

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
comm = NativeComm(unique_name("dbg"), rank, world, local, staging_bytes=16 << 20, heap_bytes=64 << 20)
cfg = GPT2Config(vocab_size=1000, n_positions=64, n_embd=256, n_layer=2, n_head=4, lm_chunk_rows=128)
batch = synthetic_batch(2, 2, 64, cfg.vocab_size, device=dev, seed=100 * rank)


def grads(mode, graph=False):
    torch.manual_seed(7)
    model = GPT2DoubleHeads(cfg).to(dev)
    comm.heap_reset()
    if mode == "local":          # no communication at all: the local gradients
        eng = FlatDataParallel(model, None, world_size=1, rank=0, lr=0.0, max_norm=0.0, bucket_mb=0.5)
    else:
        eng = FlatDataParallel(model, comm, world_size=world, rank=rank, lr=0.0, max_norm=0.0, bucket_mb=0.5,
                               zero1=False)        # plain DP: the flat buffer then holds the full averaged gradient
    if graph:
        eng.capture(batch, warmup=0)
        eng._graph.replay()
    else:
        eng.step(batch)
    torch.cuda.synchronize()
    g = eng.flat_grad.float().clone()
    names = [(n, o, p.numel()) for (n, p), o in zip(model.named_parameters(), eng._offsets)]
    eng.close()
    return g, names


g_local, names = grads("local")
want = g_local.clone()
dist.all_reduce(want)
want /= world
fail = 0
for graph in (False, True):
    g_ours, _ = grads("ours", graph)
    comm.check()
    bad = 0
    for n, o, k in names:
        a, b = g_ours[o:o + k], want[o:o + k]
        err = float((a - b).abs().max())
        ref = float(b.abs().max()) + 1e-12
        if err > 0.05 * ref + 1e-6:
            bad += 1
            if rank == 0 and bad <= 8:
                print(f"[parity] graph={graph} {n}: max err {err:.3g} vs ref max {ref:.3g}; distance to the LOCAL grad "
                      f"{float((a - g_local[o:o + k]).abs().max()):.3g}", flush=True)
    fail += bad
    if rank == 0:
        print(f"[parity] graph={graph}: parameters with wrong averaged gradients: {bad} of {len(names)}", flush=True)
t = torch.tensor([fail], device=dev)
dist.all_reduce(t)
dist.barrier()
comm.close()
dist.destroy_process_group()
sys.exit(1 if t.item() else 0)
