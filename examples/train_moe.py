"""MoE workload with relay control — BASELINE.json config 5 ("MoE DDP with relay control: 2 of 8
ranks forced idle as relays"); reference script: /root/reference/models/moe/train_moe.py
(FMoETransformerMLP(num_expert=10, d_model=1024, d_hidden=4096, top_k=1), 20 iterations).

A small transformer-ish block (dense projection + MoE MLP) is wrapped in DDP with the AdapCC comm
hook. ``--stragglers r0,r1`` makes those ranks sleep before backward every step from step 2 on, so
the coordinator's ski-rental rule excludes them: the other ranks all-reduce among themselves while
the stragglers become relays (BSP: they apply their own gradient). ``--expert_parallel`` shards the
experts over the ranks and exchanges tokens with the in-kernel dispatch/combine path.

    torchrun --nproc-per-node 8 examples/train_moe.py --stragglers 6,7 --steps 10
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adapcc_b200 import ALLREDUCE  # noqa: E402
from adapcc_b200.adapcc import AdapCC  # noqa: E402
from adapcc_b200.models.moe import MoEMLP  # noqa: E402
from adapcc_b200.parallel.ddp import wrap_ddp  # noqa: E402


class Block(nn.Module):
    def __init__(self, d, moe):
        super().__init__()
        self.norm = nn.LayerNorm(d)
        self.proj = nn.Linear(d, d)
        self.moe = moe

    def forward(self, x):
        return (x + self.moe(self.proj(self.norm(x)))).float().pow(2).mean()


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--port", default="5000")
    p.add_argument("--strategy_file", default="./strategy/moe.xml")
    p.add_argument("--logical_graph", default="./topology/logical_graph.xml")
    p.add_argument("--entry_point", type=int, default=-1)
    p.add_argument("--parallel_degree", type=int, default=4)
    p.add_argument("--profile_freq", type=int, default=0)
    p.add_argument("--num_expert", type=int, default=10)
    p.add_argument("--d_model", type=int, default=1024)
    p.add_argument("--d_hidden", type=int, default=4096)
    p.add_argument("--top_k", type=int, default=1)
    p.add_argument("--tokens", type=int, default=3200)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--stragglers", default="")
    p.add_argument("--straggle_ms", type=float, default=250.0)
    p.add_argument("--relay_mode", default="bypass", choices=["forward", "bypass"])
    p.add_argument("--algo", default="auto")
    p.add_argument("--expert_parallel", action="store_true")
    a = p.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "1234")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    stragglers = {int(x) for x in a.stragglers.split(",") if x.strip()}
    a.relay_control = True
    a.relay_threshold = 0.05
    a.heap_mb = 512 if a.expert_parallel else 0
    AdapCC.init(a, local, rank, world)
    AdapCC.setup(ALLREDUCE)
    comm = AdapCC.communicator
    exchange = None
    if a.expert_parallel and world > 1:
        from adapcc_b200.parallel.expert_parallel import ExpertExchange

        probe = MoEMLP(a.num_expert, a.d_model, a.d_hidden, a.top_k, world_size=world)
        exchange = ExpertExchange(comm.native, a.num_expert, probe.capacity(a.tokens), a.d_model)
    torch.manual_seed(0)
    moe = MoEMLP(a.num_expert, a.d_model, a.d_hidden, a.top_k, world_size=world if exchange else 1, exchange=exchange)
    model = Block(a.d_model, moe).to(dev).bfloat16()
    if exchange is not None:                       # expert weights are rank-local (fastmoe dp_comm="none")
        ignore = [n for n, q in model.named_parameters() if getattr(q, "expert_parallel", False)]
        torch.nn.parallel.DistributedDataParallel._set_params_and_buffers_to_ignore_for_model(model, ignore)
    ddp = wrap_ddp(model, comm, local, zero_copy=False)
    opt = torch.optim.SGD(ddp.parameters(), lr=1e-3)
    for i in range(a.steps):
        comm.update_relay(step=i)
        t0 = time.time()
        x = torch.rand(a.tokens, a.d_model, device=dev, dtype=torch.bfloat16)
        loss = ddp(x)
        opt.zero_grad(set_to_none=False)
        if rank in stragglers and i >= 2:
            time.sleep(a.straggle_ms / 1e3)        # this rank's first bucket arrives late
        loss.backward()
        opt.step()
        torch.cuda.synchronize()
        print("[rank %d] step %d loss %.4f active %s %.1f ms" % (rank, i, loss.item(), comm.active_gpus,
                                                                  (time.time() - t0) * 1e3), flush=True)
    comm.synchronize()
    if rank == 0:
        print("relay steps per rank are reported by each rank's communicator.stats", flush=True)
    print("[rank %d] relay_steps=%d" % (rank, comm.stats["relay_steps"]), flush=True)
    AdapCC.clear(ALLREDUCE)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
