"""Template training script — the reference's ``train_ddp.py`` on this library.

Same steps as /root/reference/train_ddp.py:30-58: init the process group, wrap the model in DDP,
``AdapCC.init`` -> ``AdapCC.setup(ALLREDUCE)`` -> ``register_comm_hook(cuda_allreduce_hook)``, then per
step ``update_relay(step)``, every ``profile_freq`` steps ``reconstruct_topology`` (re-profile the
links, re-synthesise the strategy), ``AdapCC.clear`` at the end. Launch with torchrun or
``python -m adapcc_b200.launcher``; rank ids come from RANK/LOCAL_RANK/WORLD_SIZE (OMPI_* accepted).

    torchrun --nproc-per-node 8 train_ddp.py --model vgg16 --entry_point 7 --strategy_file strategy/8.xml
"""
import argparse
import os
import time

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.optim as optim

from adapcc_b200 import ALLREDUCE
from adapcc_b200.adapcc import AdapCC
from adapcc_b200.parallel.ddp import rebuild_buckets, wrap_ddp


def env_int(a, b, d):
    return int(os.environ.get(a, os.environ.get(b, d)))


LOCAL_RANK = env_int("LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", 0)
WORLD_SIZE = env_int("WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", 1)
WORLD_RANK = env_int("RANK", "OMPI_COMM_WORLD_RANK", 0)


def build_model(name: str):
    if name == "vgg16":
        import torchvision.models as models

        return models.vgg16(), (3, 224, 224), 1000
    if name == "resnet18":
        import torchvision.models as models

        return models.resnet18(), (3, 224, 224), 1000
    if name == "mlp":
        return nn.Sequential(nn.Flatten(), nn.Linear(3 * 32 * 32, 1024), nn.ReLU(), nn.Linear(1024, 10)), (3, 32, 32), 10
    raise SystemExit(f"unknown model {name}")


def init_processes(args):
    use_cuda = args.backend == "nccl" and torch.cuda.is_available()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "1234")
    dev = torch.device("cuda", LOCAL_RANK) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(LOCAL_RANK)
        dist.init_process_group(args.backend, rank=WORLD_RANK, world_size=WORLD_SIZE, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=WORLD_RANK, world_size=WORLD_SIZE)
        args.backend = "gloo"
    model, shape, classes = build_model(args.model)
    # whole-model precision (the reference's accuracy benchmark casts the model with --fp16 / --bfp16,
    # models/image-classification/accuracy_benchmark.py:93-94,196-199)
    mdtype = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}[args.dtype]
    model = model.to(dev).to(mdtype)
    loss_fn = nn.CrossEntropyLoss()

    AdapCC.init(args, LOCAL_RANK, WORLD_RANK, WORLD_SIZE)
    AdapCC.setup(ALLREDUCE)
    ddp_model = wrap_ddp(model, AdapCC.communicator, LOCAL_RANK, bucket_cap_mb=args.bucket_cap_mb,
                         zero_copy=use_cuda and args.heap_mb > 0)
    optimizer = optim.SGD(ddp_model.parameters(), lr=0.001)

    for i in range(args.steps):
        # reconstruct BEFORE the step's heartbeat: a heartbeat sent to the old coordinator would never be answered
        # (no hook fires for it), so its controller thread would sit in the RPC until clear() gave up on the join
        if i != 0 and AdapCC.profile_freq and i % AdapCC.profile_freq == 0:
            AdapCC.reconstruct_topology(args, ALLREDUCE)
        AdapCC.communicator.update_relay(step=i)
        t0 = time.time()
        outputs = ddp_model(torch.randn(args.batch, *shape, device=dev, dtype=mdtype))
        labels = torch.randint(0, classes, [args.batch], device=dev)
        loss = loss_fn(outputs.float(), labels)
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        if i == 0 and use_cuda and args.heap_mb > 0:
            rebuild_buckets(ddp_model, AdapCC.communicator)   # DDP's one-off re-bucketing, into the heap
        if WORLD_RANK == 0:
            print("======== step %d \t loss %0.3f \t %.1f ms" % (i, loss.item(), (time.time() - t0) * 1e3), flush=True)
    AdapCC.communicator.synchronize()
    AdapCC.clear(ALLREDUCE)
    dist.destroy_process_group()


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--backend", type=str, default="nccl", choices=["nccl", "gloo"])
    parser.add_argument("--port", type=str, default="5000")
    parser.add_argument("--strategy_file", type=str, default="./strategy/strategy.xml")
    parser.add_argument("--logical_graph", type=str, default="./topology/logical_graph.xml")
    parser.add_argument("--entry_point", type=int, default=-1)
    parser.add_argument("--parallel_degree", type=int, default=4)
    parser.add_argument("--profile_freq", type=int, default=500)
    parser.add_argument("--model", type=str, default="vgg16")
    parser.add_argument("--batch", type=int, default=64)
    parser.add_argument("--steps", type=int, default=5)
    parser.add_argument("--bucket_cap_mb", type=int, default=100)
    parser.add_argument("--heap_mb", type=int, default=0, help=">0: DDP buckets live in the symmetric heap (zero-copy)")
    parser.add_argument("--wire_dtype", type=str, default=None, help="e.g. bfloat16: fp32 buckets travel as bf16")
    parser.add_argument("--dtype", type=str, default="fp32", choices=["fp32", "fp16", "bf16"],
                        help="whole-model precision (gradient buckets travel in the same dtype)")
    parser.add_argument("--algo", type=str, default="auto")
    parser.add_argument("--relay_mode", type=str, default="forward", choices=["forward", "bypass"])
    init_processes(parser.parse_args())
