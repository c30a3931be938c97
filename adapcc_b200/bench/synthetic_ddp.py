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
    ap.add_argument("--fp16-allreduce", action="store_true",
                    help="the Horovod script's flag: compress fp32 gradients on the wire (= --wire_dtype float16)")
    ap.add_argument("--backend", default="nccl", choices=["nccl", "gloo"], help="gloo: CPU run (tests)")
    ap.add_argument("--image_size", type=int, default=224)
    ap.add_argument("--num-classes", type=int, default=1000)
    a = ap.parse_args()
    if a.fp16_allreduce and not a.wire_dtype:
        a.wire_dtype = "float16"
    import torchvision.models as models

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    cuda = a.backend == "nccl" and torch.cuda.is_available()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "1234")
    if cuda:
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dev = torch.device("cpu")
        dist.init_process_group("gloo", rank=rank, world_size=world)
    sync = torch.cuda.synchronize if cuda else (lambda: None)
    model = getattr(models, a.model)(num_classes=a.num_classes).to(dev)
    if a.hook == "adapcc":
        from .. import ALLREDUCE
        from ..adapcc import AdapCC
        from ..parallel.ddp import wrap_ddp

        args = SimpleNamespace(port=5000, strategy_file="./strategy/synthetic.xml", logical_graph="./topology/lg.xml",
                               entry_point=-1, parallel_degree=4, profile_freq=0, relay_control=False,
                               wire_dtype=a.wire_dtype, backend="nccl" if cuda else "gloo")
        AdapCC.init(args, local, rank, world)
        AdapCC.setup(ALLREDUCE)
        ddp = wrap_ddp(model, AdapCC.communicator, local, zero_copy=False)
    else:
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local] if cuda else None)
    opt = torch.optim.SGD(ddp.parameters(), lr=0.01)
    data = torch.randn(a.batch_size, 3, a.image_size, a.image_size, device=dev)
    target = torch.randint(0, a.num_classes, (a.batch_size,), device=dev)

    def step():
        opt.zero_grad(set_to_none=False)
        F.cross_entropy(ddp(data), target).backward()
        opt.step()

    for _ in range(a.num_warmup_batches):
        step()
    rates = []
    for _ in range(a.num_iters):
        sync()
        t0 = time.time()
        for _ in range(a.num_batches_per_iter):
            step()
        sync()
        rates.append(a.batch_size * a.num_batches_per_iter / (time.time() - t0))
    if rank == 0:
        m = sum(rates) / len(rates)
        conf = 1.96 * (sum((r - m) ** 2 for r in rates) / len(rates)) ** 0.5       # the Horovod script's +- column
        print(f"Img/sec per GPU: {m:.1f} +-{conf:.1f}; total on {world} GPU(s): {m * world:.1f} +-{conf * world:.1f} "
              f"(hook={a.hook}, wire={a.wire_dtype or 'native'})")
    if a.hook == "adapcc":
        AdapCC.clear(ALLREDUCE)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
