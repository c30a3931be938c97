"""Straggler-gap measurement — the reference's units-test/get_wait_time.py: a recording coordinator
timestamps every rank's first-bucket arrival per step and writes ``(max - min) * heter_alpha`` per
step to a CSV (/root/reference/units-test/get_wait_time.py:50-62,103; results
wait_time_{homo,heter}_bc128.csv). Here the production coordinator already logs arrivals
(``Coordinator.arrival_log``), so the script only trains a DDP model with the hook and dumps the gaps.

    torchrun --nproc-per-node 8 -m adapcc_b200.bench.wait_time --steps 200 --out wait_time.csv [--heter_alpha 2.7]
"""
import argparse
import os
from types import SimpleNamespace

import torch
import torch.distributed as dist


def gaps_from_log(arrival_log, alpha: float = 1.0, skip: int = 2):
    """[(step, gap_seconds * alpha)] for every step where at least two ranks reported."""
    out = []
    for step in sorted(arrival_log):
        ts = [t for _, t in arrival_log[step]]
        if step >= skip and len(ts) >= 2:
            out.append((step, (max(ts) - min(ts)) * alpha))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--heter_alpha", type=float, default=1.0)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--out", default="wait_time.csv")
    ap.add_argument("--backend", default="nccl", choices=["nccl", "gloo"])
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    cuda = a.backend == "nccl" and torch.cuda.is_available()
    dev = torch.device("cuda", local) if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from .. import ALLREDUCE
    from ..adapcc import AdapCC

    args = SimpleNamespace(port=5000, strategy_file="./strategy/wait.xml", logical_graph="./topology/lg.xml",
                           entry_point=-1, parallel_degree=4, profile_freq=0, relay_control=True, coordinator_process=False,
                           relay_threshold=10.0,                      # never exclude anybody: we only measure
                           backend="nccl" if cuda else "gloo")
    AdapCC.init(args, local, rank, world)
    AdapCC.setup(ALLREDUCE)
    comm = AdapCC.communicator
    model = torch.nn.Sequential(torch.nn.Linear(1024, 4096), torch.nn.ReLU(), torch.nn.Linear(4096, 1024)).to(dev)
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local] if cuda else None)
    ddp.register_comm_hook(state=None, hook=comm.cuda_allreduce_hook)
    opt = torch.optim.SGD(ddp.parameters(), lr=1e-3)
    for step in range(a.steps):
        comm.update_relay(step)
        loss = ddp(torch.randn(a.batch, 1024, device=dev)).pow(2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
    comm.synchronize()
    if rank == 0:
        rows = gaps_from_log(comm.coordinator.arrival_log, a.heter_alpha)
        with open(a.out, "w") as f:
            f.writelines(f"{s},{g:.6f}\n" for s, g in rows)
        if rows:
            vals = sorted(g for _, g in rows)
            print(f"steps {len(vals)} mean {sum(vals) / len(vals) * 1e3:.3f} ms median {vals[len(vals) // 2] * 1e3:.3f} ms "
                  f"p95 {vals[int(len(vals) * 0.95)] * 1e3:.3f} ms (alpha {a.heter_alpha})")
    AdapCC.clear(ALLREDUCE)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
