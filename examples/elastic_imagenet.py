"""Elastic data-parallel training with checkpoint / resume — the reference's torchelastic ImageNet
example (/root/reference/models/image-classification/main_elastic.py, launch_elastic.sh:
``torchrun --nnodes=1:3 --max_restarts=3 --rdzv_backend=c10d``) on synthetic data and with the AdapCC
comm hook. On (re)start every worker loads its newest local checkpoint, the newest one in the job is
broadcast to everybody (a restarted or newly joined worker may be behind), training resumes from that
epoch; checkpoints are written atomically by local rank 0 after every epoch.

    torchrun --nnodes=1:3 --nproc-per-node 8 --max_restarts 3 --rdzv_backend c10d \
        --rdzv_endpoint 127.0.0.1:29400 examples/elastic_imagenet.py --epochs 3
"""
import argparse
import os
import sys
from types import SimpleNamespace

import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adapcc_b200 import ALLREDUCE  # noqa: E402
from adapcc_b200.adapcc import AdapCC  # noqa: E402
from adapcc_b200.parallel.ddp import wrap_ddp  # noqa: E402
from adapcc_b200.utils.checkpoint import State, broadcast_newest  # noqa: E402
from adapcc_b200.utils.meters import AverageMeter, ProgressMeter  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--arch", default="resnet18")
    p.add_argument("--epochs", type=int, default=2)
    p.add_argument("--steps_per_epoch", type=int, default=20)
    p.add_argument("--batch", type=int, default=32)
    p.add_argument("--checkpoint", default="./checkpoint.pt")
    p.add_argument("--backend", default="nccl", choices=["nccl", "gloo"])
    a = p.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    cuda = a.backend == "nccl" and torch.cuda.is_available()
    dev = torch.device("cuda", local) if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(local)
    dist.init_process_group(a.backend if cuda else "gloo")
    import torchvision.models as models

    model = getattr(models, a.arch)(num_classes=100).to(dev)
    opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9)
    state = State(model, opt)
    state.load(a.checkpoint)                                  # may be absent or stale on this worker
    gloo = dist.new_group(backend="gloo") if cuda else None   # snapshots travel over a CPU group
    src = broadcast_newest(state, group=gloo)
    if rank == 0:
        print(f"=> resuming from epoch {state.epoch + 1} (newest checkpoint held by rank {src})", flush=True)
    args = SimpleNamespace(port=5000, strategy_file="./strategy/elastic.xml", logical_graph="./topology/lg.xml",
                           entry_point=-1, parallel_degree=4, profile_freq=0, relay_control=world > 1,
                           backend="nccl" if cuda else "gloo")
    AdapCC.init(args, local, rank, world)
    AdapCC.setup(ALLREDUCE)
    ddp = wrap_ddp(model, AdapCC.communicator, local, zero_copy=False)
    step = 0
    for epoch in range(state.epoch + 1, a.epochs):
        losses = AverageMeter("Loss", ":.4e")
        prog = ProgressMeter(a.steps_per_epoch, [losses], prefix=f"Epoch: [{epoch}]")
        for i in range(a.steps_per_epoch):
            AdapCC.communicator.update_relay(step)
            step += 1
            x = torch.randn(a.batch, 3, 64, 64, device=dev)
            y = torch.randint(0, 100, (a.batch,), device=dev)
            loss = F.cross_entropy(ddp(x), y)
            opt.zero_grad(set_to_none=False)
            loss.backward()
            opt.step()
            losses.update(loss.item(), a.batch)
            if rank == 0 and i % 10 == 0:
                prog.display(i)
        state.epoch, state.step = epoch, step
        if local == 0:
            state.save(a.checkpoint)                          # atomic: tmp + rename
    AdapCC.communicator.synchronize()
    AdapCC.clear(ALLREDUCE)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
