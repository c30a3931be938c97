"""torchrun debug worker: one eager step of the flat engine with ADAPCC_ENGINE_DEBUG=1 (which parameter launches which
bucket, double 'ready' reports), then per-parameter comparison of the all-reduced gradients against an NCCL all-reduce of
the same local gradients."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ["ADAPCC_ENGINE_DEBUG"] = "1"
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


def grads(mode):
    torch.manual_seed(7)
    model = GPT2DoubleHeads(cfg).to(dev)
    comm.heap_reset()
    if mode == "local":          # no communication at all: the local gradients
        eng = FlatDataParallel(model, None, world_size=1, rank=0, lr=0.0, max_norm=0.0, bucket_mb=0.5)
    else:
        eng = FlatDataParallel(model, comm, world_size=world, rank=rank, lr=0.0, max_norm=0.0, bucket_mb=0.5, zero1=False)
    if rank != 0:
        eng._debug = False
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
g_ours, _ = grads("ours")
comm.check()
if rank == 0:
    bad = 0
    for n, o, k in names:
        a, b = g_ours[o:o + k], want[o:o + k]
        err = float((a - b).abs().max())
        ref = float(b.abs().max()) + 1e-12
        if err > 0.05 * ref + 1e-6:
            bad += 1
            la = float((g_ours[o:o + k] - g_local[o:o + k]).abs().max())
            print(f"[dbg] {n}: all-reduced grad differs from the NCCL average: max err {err:.3g} (ref max {ref:.3g}); "
                  f"distance to the LOCAL grad {la:.3g}", flush=True)
    print(f"[dbg] parameters with wrong averaged gradients: {bad} of {len(names)}", flush=True)
dist.barrier()
comm.close()
dist.destroy_process_group()
