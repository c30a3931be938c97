"""Synthetic image-model DDP throughput (img/s) — the reference's Horovod benchmark
(/root/reference/nccl-perf/pytorch_synthetic.py:38-117) on torch DDP, with or without the AdapCC hook.

    torchrun --nproc-per-node 8 -m adapcc_b200.bench.synthetic_ddp --model resnet50 --hook adapcc
"""
import argparse
import os
import time
from types import SimpleNamespace

import torch
import torch.distributed as dist
import torch.nn.functional as F


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="resnet50")
    ap.add_argument("--batch-size", type=int, default=32)
    ap.add_argument("--num-warmup-batches", type=int, default=5)
    ap.add_argument("--num-batches-per-iter", type=int, default=10)
    ap.add_argument("--num-iters", type=int, default=5)
    ap.add_argument("--hook", default="adapcc", choices=["adapcc", "nccl"])
    ap.add_argument("--wire_dtype", default=None)
    a = ap.parse_args()
    import torchvision.models as models

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    model = getattr(models, a.model)().to(dev)
    if a.hook == "adapcc":
        from .. import ALLREDUCE
        from ..adapcc import AdapCC
        from ..parallel.ddp import wrap_ddp

        args = SimpleNamespace(port=5000, strategy_file="./strategy/synthetic.xml", logical_graph="./topology/lg.xml",
                               entry_point=-1, parallel_degree=4, profile_freq=0, relay_control=False,
                               wire_dtype=a.wire_dtype)
        AdapCC.init(args, local, rank, world)
        AdapCC.setup(ALLREDUCE)
        ddp = wrap_ddp(model, AdapCC.communicator, local, zero_copy=False)
    else:
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
    opt = torch.optim.SGD(ddp.parameters(), lr=0.01)
    data = torch.randn(a.batch_size, 3, 224, 224, device=dev)
    target = torch.randint(0, 1000, (a.batch_size,), device=dev)

    def step():
        opt.zero_grad(set_to_none=False)
        F.cross_entropy(ddp(data), target).backward()
        opt.step()

    for _ in range(a.num_warmup_batches):
        step()
    rates = []
    for _ in range(a.num_iters):
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(a.num_batches_per_iter):
            step()
        torch.cuda.synchronize()
        rates.append(a.batch_size * a.num_batches_per_iter / (time.time() - t0))
    if rank == 0:
        m = sum(rates) / len(rates)
        print(f"Img/sec per GPU: {m:.1f}; total on {world} GPU(s): {m * world:.1f} (hook={a.hook})")
    if a.hook == "adapcc":
        AdapCC.clear(ALLREDUCE)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
