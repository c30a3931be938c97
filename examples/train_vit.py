"""ViT DDP with periodic reconstruct_topology — BASELINE.json config 4 ("ViT-B/16 DDP with
reconstruct_topology every 500 steps (profiling path)"); reference script:
/root/reference/models/vit/train_vit.py (synthetic batches, SGD, prints step time).

    torchrun --nproc-per-node 8 examples/train_vit.py --entry_point 7 --profile_freq 500 --steps 1001
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adapcc_b200 import ALLREDUCE  # noqa: E402
from adapcc_b200.adapcc import AdapCC  # noqa: E402
from adapcc_b200.models.vit import ViT, ViTConfig  # noqa: E402
from adapcc_b200.parallel.ddp import wrap_ddp  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--port", default="5000")
    p.add_argument("--strategy_file", default="./strategy/vit.xml")
    p.add_argument("--logical_graph", default="./topology/logical_graph.xml")
    p.add_argument("--entry_point", type=int, default=7)
    p.add_argument("--parallel_degree", type=int, default=4)
    p.add_argument("--profile_freq", type=int, default=500)
    p.add_argument("--shape", default="b16", choices=["b16", "reference", "tiny"])
    p.add_argument("--batch", type=int, default=256)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--wire_dtype", default=None)
    a = p.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "1234")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    cfg = getattr(ViTConfig, a.shape)()
    model = ViT(cfg).to(dev).bfloat16()
    AdapCC.init(a, local, rank, world)
    AdapCC.setup(ALLREDUCE)
    ddp = wrap_ddp(model, AdapCC.communicator, local, zero_copy=False)
    opt = torch.optim.SGD(ddp.parameters(), lr=1e-3)
    for i in range(a.steps):
        if i and AdapCC.profile_freq and i % AdapCC.profile_freq == 0:
            t0 = time.time()
            AdapCC.reconstruct_topology(a, ALLREDUCE)     # re-profile links, re-synthesise, new contexts
            if rank == 0:
                print("reconstruct_topology: %.1f ms" % ((time.time() - t0) * 1e3), flush=True)
        AdapCC.communicator.update_relay(step=i)          # after the reconstruct: heartbeat goes to the LIVE coordinator
        t0 = time.time()
        x = torch.randn(a.batch, 3, cfg.image_size, cfg.image_size, device=dev, dtype=torch.bfloat16)
        y = torch.randint(0, cfg.num_classes, (a.batch,), device=dev)
        loss = ddp(x, y)[0]
        opt.zero_grad(set_to_none=False)
        loss.backward()
        opt.step()
        torch.cuda.synchronize()
        if rank == 0:
            print("step %d loss %.3f time %.1f ms" % (i, loss.item(), (time.time() - t0) * 1e3), flush=True)
    AdapCC.communicator.synchronize()
    AdapCC.clear(ALLREDUCE)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
