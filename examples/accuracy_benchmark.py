"""ImageNet-style accuracy / precision benchmark: torchvision model, SGD + momentum + StepLR(30, 0.1), Acc@1 / Acc@5
meters, validation, checkpoint / resume, early ``--stop``, whole-model ``--fp16`` / ``--bfp16``, gradient-noise-scale
probe — the reference's /root/reference/models/image-classification/accuracy_benchmark.py (its logs
``accuracy_*.txt`` / ``resnet18_{fp32,fp16,bfp16}.txt`` are the per-print Acc@1 columns, extracted by
``process_log.py``; ``gns-split-all.txt`` the "mean gns:" lines, extracted by ``process_gns.py`` — here
``tools/process_log.py`` does both).

Differences: gradients travel through this library (torch DDP + ``cuda_allreduce_hook``; ``--wire_dtype bfloat16``
additionally compresses fp32 buckets on the wire, which is the precision question the benchmark asks of a
communication library), ``--dummy`` data is *learnable* (class templates + noise, generated per index — the reference's
``FakeData`` is pure noise, so its dummy accuracy stays at chance), the GNS probe is live (the reference has it
commented out), validation partitions the set exactly (no padded duplicates), and everything runs under
``--backend gloo`` on CPU for tests.

    torchrun --nproc-per-node 8 examples/accuracy_benchmark.py /data/imagenet -a resnet18 -b 256 --bfp16
    torchrun --nproc-per-node 2 examples/accuracy_benchmark.py --dummy --backend gloo -a resnet18 --image_size 32 \
        --classes 10 --dummy_size 512 --epochs 2 -b 32
"""
import argparse
import os
import shutil
import sys
import time
from types import SimpleNamespace

import torch
import torch.distributed as dist
import torch.nn as nn
from torch.optim.lr_scheduler import StepLR

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adapcc_b200 import ALLREDUCE  # noqa: E402
from adapcc_b200.adapcc import AdapCC  # noqa: E402
from adapcc_b200.parallel.ddp import wrap_ddp  # noqa: E402
from adapcc_b200.utils.gns import GNSProbe, grad_sq_norm  # noqa: E402
from adapcc_b200.utils.meters import AverageMeter, ProgressMeter  # noqa: E402


class TemplateImages(torch.utils.data.Dataset):
    """Learnable stand-in for ImageNet: class c = a fixed low-resolution random template, upsampled, plus per-sample
    noise; sample i is a pure function of (seed, i)."""

    def __init__(self, n: int, classes: int, size: int, seed: int = 0, noise: float = 1.0):
        g = torch.Generator().manual_seed(seed)
        self.templates = torch.nn.functional.interpolate(torch.randn(classes, 3, 8, 8, generator=g), size=(size, size),
                                                         mode="bilinear", align_corners=False)
        self.n, self.classes, self.seed, self.noise = n, classes, seed, noise

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 1_000_003 + i)
        y = int(torch.randint(0, self.classes, (1,), generator=g))
        return self.templates[y] + self.noise * torch.randn(self.templates.shape[1:], generator=g), y


class ExactPartition(torch.utils.data.Sampler):
    """rank r scores indices r, r + world, …: every sample exactly once across the job."""

    def __init__(self, n, rank, world):
        self.idx = list(range(rank, n, world))

    def __iter__(self):
        return iter(self.idx)

    def __len__(self):
        return len(self.idx)


def accuracy(output, target, topk=(1,)):
    """Percentage of samples whose label is among the k highest logits."""
    maxk = min(max(topk), output.shape[1])
    pred = output.float().topk(maxk, 1).indices
    hit = pred == target[:, None]
    return [hit[:, :min(k, maxk)].any(1).float().mean() * 100.0 for k in topk]


def reduce_meter(m: AverageMeter, dev):
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([m.sum, m.count], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        m.sum, m.count = float(t[0]), int(t[1])
        m.avg = m.sum / max(1, m.count)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("data", nargs="?", default="imagenet", help="dataset root with train/ and val/ class folders")
    p.add_argument("-a", "--arch", default="resnet18")
    p.add_argument("-j", "--workers", type=int, default=4)
    p.add_argument("--epochs", type=int, default=90)
    p.add_argument("--start-epoch", type=int, default=0)
    p.add_argument("-b", "--batch-size", type=int, default=256, help="per process")
    p.add_argument("--lr", "--learning-rate", type=float, default=0.1, dest="lr")
    p.add_argument("--momentum", type=float, default=0.9)
    p.add_argument("--wd", "--weight-decay", type=float, default=1e-4, dest="weight_decay")
    p.add_argument("-p", "--print-freq", type=int, default=10)
    p.add_argument("--resume", default="")
    p.add_argument("-e", "--evaluate", action="store_true")
    p.add_argument("--seed", type=int, default=None)
    p.add_argument("--dummy", action="store_true", help="learnable synthetic data instead of an image folder")
    p.add_argument("--dummy_size", type=int, default=4096)
    p.add_argument("--classes", type=int, default=1000)
    p.add_argument("--image_size", type=int, default=224)
    p.add_argument("--stop", type=int, default=-1, help="leave every epoch after this many steps")
    p.add_argument("--fp16", action="store_true", help="whole model + inputs in fp16")
    p.add_argument("--bfp16", action="store_true", help="whole model + inputs in bf16")
    p.add_argument("--gns_freq", type=int, default=0, help=">0: print 'mean gns:' every this many steps (one extra local backward)")
    p.add_argument("--checkpoint", default="checkpoint.pth.tar")
    p.add_argument("--backend", default="nccl", choices=["nccl", "gloo"])
    p.add_argument("--wire_dtype", default=None, help="e.g. bfloat16: fp32 gradient buckets travel compressed")
    p.add_argument("--algo", default="auto")
    p.add_argument("--port", default="5000")
    p.add_argument("--strategy_file", default="./strategy/accuracy.xml")
    p.add_argument("--logical_graph", default="./topology/logical_graph.xml")
    p.add_argument("--entry_point", type=int, default=-1)
    p.add_argument("--parallel_degree", type=int, default=4)
    p.add_argument("--profile_freq", type=int, default=0)
    a = p.parse_args()

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    cuda = a.backend == "nccl" and torch.cuda.is_available()
    dev = torch.device("cuda", local) if cuda else torch.device("cpu")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "1234")
    if cuda:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        a.backend = "gloo"
        dist.init_process_group("gloo", rank=rank, world_size=world)
    if a.seed is not None:
        torch.manual_seed(a.seed)
    import torchvision.models as models

    if rank == 0:
        print(f"=> creating model '{a.arch}'", flush=True)
    model = getattr(models, a.arch)(num_classes=a.classes).to(dev)
    dtype = torch.float16 if a.fp16 else torch.bfloat16 if a.bfp16 else torch.float32
    model = model.to(dtype)
    criterion = nn.CrossEntropyLoss()
    optimizer = torch.optim.SGD(model.parameters(), a.lr, momentum=a.momentum, weight_decay=a.weight_decay)
    scheduler = StepLR(optimizer, step_size=30, gamma=0.1)
    best_acc1 = 0.0
    if a.resume and os.path.isfile(a.resume):
        ck = torch.load(a.resume, map_location="cpu", weights_only=False)
        a.start_epoch, best_acc1 = ck["epoch"], float(ck["best_acc1"])
        model.load_state_dict(ck["state_dict"])
        optimizer.load_state_dict(ck["optimizer"])
        scheduler.load_state_dict(ck["scheduler"])
        if rank == 0:
            print(f"=> loaded checkpoint '{a.resume}' (epoch {ck['epoch']})", flush=True)

    # ---- data -------------------------------------------------------------------------------------------------
    if a.dummy:
        if rank == 0:
            print("=> Dummy data is used!", flush=True)
        train_set = TemplateImages(a.dummy_size, a.classes, a.image_size, seed=1)
        val_set = TemplateImages(max(64, a.dummy_size // 8), a.classes, a.image_size, seed=1)
        val_set.seed = 2                                               # same templates, unseen noise
        workers = 0
    else:
        import torchvision.datasets as datasets
        import torchvision.transforms as T

        norm = T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
        train_set = datasets.ImageFolder(os.path.join(a.data, "train"), T.Compose(
            [T.RandomResizedCrop(a.image_size), T.RandomHorizontalFlip(), T.ToTensor(), norm]))
        val_set = datasets.ImageFolder(os.path.join(a.data, "val"), T.Compose(
            [T.Resize(int(a.image_size * 256 / 224)), T.CenterCrop(a.image_size), T.ToTensor(), norm]))
        workers = a.workers
    sampler = torch.utils.data.distributed.DistributedSampler(train_set, world, rank, drop_last=True) if world > 1 else None
    train_loader = torch.utils.data.DataLoader(train_set, batch_size=a.batch_size, shuffle=sampler is None, sampler=sampler,
                                               num_workers=workers, pin_memory=cuda, drop_last=True)
    val_loader = torch.utils.data.DataLoader(val_set, batch_size=a.batch_size, sampler=ExactPartition(len(val_set), rank, world),
                                             num_workers=workers, pin_memory=cuda)

    # ---- communication ----------------------------------------------------------------------------------------
    args = SimpleNamespace(port=a.port, strategy_file=a.strategy_file, logical_graph=a.logical_graph, entry_point=a.entry_point,
                           parallel_degree=a.parallel_degree, profile_freq=a.profile_freq, relay_control=False,
                           backend=a.backend, wire_dtype=a.wire_dtype, algo=a.algo)
    AdapCC.init(args, local, rank, world)
    AdapCC.setup(ALLREDUCE)
    ddp = wrap_ddp(model, AdapCC.communicator, local, zero_copy=False) if world > 1 else model
    probe = GNSProbe(world, a.batch_size) if (a.gns_freq > 0 and world > 1) else None

    def validate():
        top1, top5, losses = AverageMeter("Acc@1", ":6.2f"), AverageMeter("Acc@5", ":6.2f"), AverageMeter("Loss", ":.4e")
        model.eval()
        with torch.no_grad():
            for images, target in val_loader:
                images, target = images.to(dev, dtype, non_blocking=True), target.to(dev, non_blocking=True)
                out = model(images)
                a1, a5 = accuracy(out, target, (1, 5))
                n = images.shape[0]
                losses.update(float(criterion(out.float(), target)), n)
                top1.update(float(a1), n)
                top5.update(float(a5), n)
        for m in (top1, top5, losses):
            reduce_meter(m, dev)
        model.train()
        if rank == 0:
            print(f" *   Acc@1 {top1.avg:.3f} Acc@5 {top5.avg:.3f} Loss {losses.avg:.4f} ({top1.count} samples)", flush=True)
        return top1.avg

    if a.evaluate:
        validate()
    else:
        for epoch in range(a.start_epoch, a.epochs):
            if sampler is not None:
                sampler.set_epoch(epoch)
            bt, dt = AverageMeter("Time", ":6.3f"), AverageMeter("Data", ":6.3f")
            losses, top1, top5 = AverageMeter("Loss", ":.4e"), AverageMeter("Acc@1", ":6.2f"), AverageMeter("Acc@5", ":6.2f")
            progress = ProgressMeter(len(train_loader), [bt, dt, losses, top1, top5], prefix=f"Epoch: [{epoch}]")
            end = time.time()
            for i, (images, target) in enumerate(train_loader):
                dt.update(time.time() - end)
                images, target = images.to(dev, dtype, non_blocking=True), target.to(dev, non_blocking=True)
                local_sq = None
                if probe is not None and i % a.gns_freq == 0:
                    with ddp.no_sync():                                   # local gradient (batch b) before any reduction
                        optimizer.zero_grad()
                        criterion(ddp(images).float(), target).backward()
                        local_sq = grad_sq_norm(model.parameters())
                out = ddp(images)
                loss = criterion(out.float(), target)
                a1, a5 = accuracy(out, target, (1, 5))
                optimizer.zero_grad()
                loss.backward()
                if local_sq is not None:
                    t = local_sq.detach().to(dev, torch.float32).reshape(1)
                    dist.all_reduce(t)
                    gns = probe.update(float(t) / world, float(grad_sq_norm(model.parameters())))
                    if rank == 0:
                        print("mean gns: %f" % gns, flush=True)
                optimizer.step()
                n = images.shape[0]
                losses.update(float(loss.detach()), n)
                top1.update(float(a1), n)
                top5.update(float(a5), n)
                bt.update(time.time() - end)
                end = time.time()
                if i % a.print_freq == 0 and rank == 0:
                    progress.display(i + 1)
                if i == a.stop:
                    break
            acc1 = validate()
            scheduler.step()
            is_best, best_acc1 = acc1 > best_acc1, max(acc1, best_acc1)
            if rank == 0 and a.checkpoint:
                torch.save({"epoch": epoch + 1, "arch": a.arch, "state_dict": model.state_dict(), "best_acc1": best_acc1,
                            "optimizer": optimizer.state_dict(), "scheduler": scheduler.state_dict()}, a.checkpoint + ".tmp")
                os.replace(a.checkpoint + ".tmp", a.checkpoint)
                if is_best:
                    shutil.copyfile(a.checkpoint, os.path.join(os.path.dirname(a.checkpoint) or ".", "model_best.pth.tar"))
    AdapCC.communicator.synchronize()
    AdapCC.clear(ALLREDUCE)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
