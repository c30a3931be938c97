"""torchrun debug worker: GPT-2 small, flat engine, plain data parallelism: after a few graph-replayed steps, which
parameters / gradients differ between the ranks?"""
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
cfg = GPT2Config()
torch.manual_seed(1234)
model = GPT2DoubleHeads(cfg).to(dev)
comm = NativeComm(unique_name("rep"), rank, world, local, staging_bytes=64 << 20, heap_bytes=(model.num_parameters() * 2 >> 20 << 20) + (128 << 20))
eng = FlatDataParallel(model, comm, world_size=world, rank=rank, lr=6.25e-5, max_norm=1.0, zero1=False)
batch = synthetic_batch(4, 2, 1024, cfg.vocab_size, device=dev, seed=1000 * rank)
names = [(n, o, p.numel()) for (n, p), o in zip(model.named_parameters(), eng._offsets)]


def report(tag, t):
    hi, lo = t.float().clone(), t.float().clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    d = (hi - lo)
    bad = [(n, float(d[o:o + k].max()), float((d[o:o + k] > 0).float().mean())) for n, o, k in names if float(d[o:o + k].max()) > 0]
    if rank == 0:
        print(f"[rep] {tag}: {len(bad)} of {len(names)} tensors differ between the ranks", flush=True)
        for n, m, f in bad[:12]:
            print(f"[rep]    {n}: max spread {m:.3g}, fraction of elements {f:.3g}", flush=True)


for mode in ("eager", "graph"):
    if mode == "graph":
        eng.capture(batch, warmup=0)
    for i in range(3):
        if mode == "graph":
            eng._graph.replay()
        else:
            eng.step(batch)
        torch.cuda.synchronize()
        comm.check()
        report(f"{mode} step {i} gradients", eng.flat_grad)
        report(f"{mode} step {i} parameters", eng.flat_param)
dist.barrier()
comm.close()
dist.destroy_process_group()
