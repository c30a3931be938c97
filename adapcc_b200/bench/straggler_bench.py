"""Straggler tolerance benchmark — the reason AdapCC exists (relay control; BASELINE.json
config 5 "2 of 8 ranks forced idle as relays").

Every rank trains the same DDP model; ranks listed in ``--stragglers`` sleep ``--straggle_ms`` before
their backward from step 2 on (a slow worker: its first gradient bucket is late). Two arms:

  nccl   : stock DDP — every all-reduce waits for the slowest rank, the whole job runs at its pace;
  adapcc : cuda_allreduce_hook + coordinator — the first-ready rank runs the ski-rental rule, the
           stragglers are declared relays for the step, the others all-reduce among themselves.

Reported: mean step time of the NON-straggler ranks (max over them), relay steps, RPC latency.

    torchrun --nproc-per-node 8 -m adapcc_b200.bench.straggler_bench --stragglers 6,7 --straggle_ms 100
"""
import argparse
import json
import os
import time
from types import SimpleNamespace

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", default="both", choices=["both", "nccl", "adapcc"])
    ap.add_argument("--stragglers", default="")
    ap.add_argument("--straggle_ms", type=float, default=100.0)
    ap.add_argument("--steps", type=int, default=14)
    ap.add_argument("--hidden", type=int, default=4096)
    ap.add_argument("--layers", type=int, default=6)
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--relay_threshold", type=float, default=0.02)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    stragglers = {int(x) for x in a.stragglers.split(",") if x.strip()}
    results = {}

    def build():
        torch.manual_seed(0)
        layers = []
        for _ in range(a.layers):
            layers += [torch.nn.Linear(a.hidden, a.hidden), torch.nn.GELU()]
        return torch.nn.Sequential(*layers).to(dev).bfloat16()

    def run(arm):
        model = build()
        comm = None
        if arm == "adapcc":
            from .. import ALLREDUCE
            from ..adapcc import AdapCC
            from ..parallel.ddp import wrap_ddp

            work = os.path.join("gpurun_out", "straggler_work")
            args = SimpleNamespace(port=5000, strategy_file=os.path.join(work, "s.xml"),
                                   logical_graph=os.path.join(work, "lg.xml"), entry_point=-1, parallel_degree=4,
                                   profile_freq=0, work_dir=work, relay_control=True, relay_mode="bypass",
                                   relay_threshold=a.relay_threshold, coordinator_port=50071)
            AdapCC.init(args, local, rank, world)
            AdapCC.setup(ALLREDUCE)
            comm = AdapCC.communicator
            ddp = wrap_ddp(model, comm, local, zero_copy=False)
        else:
            ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
        opt = torch.optim.SGD(ddp.parameters(), lr=1e-3)
        x = torch.randn(a.batch, a.hidden, device=dev, dtype=torch.bfloat16)
        times = []
        for step in range(a.steps):
            if comm is not None:
                comm.update_relay(step)
            torch.cuda.synchronize()
            t0 = time.time()
            loss = ddp(x).float().pow(2).mean()
            opt.zero_grad(set_to_none=False)
            if rank in stragglers and step >= 2:
                time.sleep(a.straggle_ms / 1e3)
            loss.backward()
            opt.step()
            torch.cuda.synchronize()
            times.append(time.time() - t0)
        steady = times[4:]
        mine = sum(steady) / len(steady)
        t = torch.tensor([mine if rank not in stragglers else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        info = {"active_rank_step_ms": t.item() * 1e3}
        if comm is not None:
            comm.synchronize()
            rs = torch.tensor([float(comm.stats["relay_steps"])], device=dev)
            dist.all_reduce(rs)
            rpc = comm.stats["hook_rpc_s"][4:]
            info.update({"relay_steps_total": int(rs.item()),
                         "hook_rpc_ms_median": sorted(rpc)[len(rpc) // 2] * 1e3 if rpc else None})
            from .. import ALLREDUCE
            from ..adapcc import AdapCC

            AdapCC.clear(ALLREDUCE)
        dist.barrier()
        return info

    arms = ["nccl", "adapcc"] if a.arm == "both" else [a.arm]
    for arm in arms:
        results[arm] = run(arm)
        if rank == 0:
            print(f"[straggler] {arm}: {json.dumps(results[arm])}", flush=True)
    if rank == 0:
        summary = {"world": world, "stragglers": sorted(stragglers), "straggle_ms": a.straggle_ms,
                   "model": f"{a.layers}x Linear({a.hidden}) bf16, batch {a.batch}", **results}
        if "nccl" in results and "adapcc" in results:
            summary["speedup_active_ranks"] = results["nccl"]["active_rank_step_ms"] / results["adapcc"]["active_rank_step_ms"]
        print("[straggler] " + json.dumps(summary), flush=True)
        if a.out:
            with open(a.out, "w") as f:
                json.dump(summary, f, indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
