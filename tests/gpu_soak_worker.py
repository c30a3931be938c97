"""Soak test of the device-side flag protocol (torchrun worker): one CUDA graph of ~60 mixed collectives — LL, one-shot,
two-shot, NVLS, strategy trees (staged and in place), reduce, broadcast, all-to-all, on staged and heap tensors, over the
full world and over active SUBSETS — replayed until >= 10^4 collectives have run back to back with no host
synchronisation in between. Every replay re-initialises every input on the device and compares every result with its
expected value (computed once on the host from the ranks' seeds); the running maximum error stays on the device and is
read at the end. ``ADAPCC_TIMEOUT_MS`` is armed, so a protocol bug shows up as a latched error word, not a hang.

    torchrun --nproc-per-node 8 tests/gpu_soak_worker.py [--ops 10000]
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adapcc_b200.constants import ALLREDUCE, BOARDCAST  # noqa: E402
from adapcc_b200.runtime.native import NativeComm  # noqa: E402
from adapcc_b200.runtime.rendezvous import unique_name  # noqa: E402


def gen(rank, n, dtype, seed):
    g = torch.Generator(device="cpu").manual_seed(seed * 1000 + rank)
    return (torch.randn(n, generator=g, dtype=torch.float32) * 2).to(dtype)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", type=int, default=10000)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    os.environ.setdefault("ADAPCC_TIMEOUT_MS", "20000")
    comm = NativeComm(unique_name("soak"), rank, world, local, staging_bytes=64 << 20, heap_bytes=256 << 20)
    allr = list(range(world))
    orders = [allr, list(reversed(allr))]
    comm.load_strategy("<trees>" + "".join(
        "<root id='%d' ip='h'>%s</root>" % (o[0], "".join("<gpu id='%d' ip='h'>" % r for r in o[1:]) + "</gpu>" * (world - 1))
        for o in orders) + "</trees>")
    subsets = [allr]
    if world >= 3:
        subsets += [[0, world - 1], list(range(1, world)), list(range(0, world, 2))]
    algos = ["one_shot", "two_shot"] + (["nvls"] if comm.multicast else [])

    # ---- the op list: (kind, params); sizes from tiny to 8 MB, odd tails included ----------------------------------
    plan = []
    seed = 0
    sizes = [3, 1000, 4099, 70001, (1 << 20) + 5, 2 << 20]
    for rnd in range(2):
        for n in sizes:
            for algo in algos:
                plan.append(("allreduce", dict(n=n, algo=algo, dtype=torch.float32, active=allr, heap=False)))
            plan.append(("allreduce", dict(n=n, algo="two_shot", dtype=torch.bfloat16, active=subsets[(rnd + n) % len(subsets)],
                                           heap=False)))
        plan.append(("allreduce", dict(n=1 << 20, algo="auto", dtype=torch.bfloat16, active=allr, heap=True)))
        plan.append(("allreduce", dict(n=(1 << 18) + 3, algo="auto", dtype=torch.float32, active=allr, heap=True)))
        plan.append(("allreduce", dict(n=777, algo="auto", dtype=torch.float32, active=allr, heap=False)))      # LL
        plan.append(("allreduce", dict(n=4096, algo="auto", dtype=torch.bfloat16, active=allr, heap=False)))    # LL
        plan.append(("tree", dict(n=70001, dtype=torch.float32, active=allr, heap=False, chunk=8192)))
        plan.append(("tree", dict(n=1 << 18, dtype=torch.float32, active=allr, heap=True, chunk=65536)))
        if world >= 3:
            plan.append(("tree", dict(n=50001, dtype=torch.float32, active=[0, world - 1], heap=False, chunk=8192)))
        plan.append(("reduce", dict(n=70001, root=world - 1, dtype=torch.float32)))
        plan.append(("broadcast", dict(n=(1 << 19) + 1, root=rnd % world, dtype=torch.bfloat16, heap=False)))
        plan.append(("broadcast", dict(n=1 << 19, root=(rnd + 1) % world, dtype=torch.float32, heap=True)))
        plan.append(("alltoall", dict(per=4099, dtype=torch.float32)))

    ops = []
    for kind, p in plan:
        seed += 1
        dtype = p["dtype"]
        n = p.get("n", p.get("per", 0) * world)
        src = gen(rank, n, dtype, seed).to(dev)
        x = comm.symm_empty(n, dtype) if p.get("heap") else torch.empty(n, dtype=dtype, device=dev)
        mine = src.float().cpu()
        if kind in ("allreduce", "tree"):
            act = p["active"]
            want = torch.stack([gen(r, n, dtype, seed).float() for r in act]).sum(0) if rank in act else mine
        elif kind == "reduce":
            want = torch.stack([gen(r, n, dtype, seed).float() for r in allr]).sum(0) if rank == p["root"] else mine
        elif kind == "broadcast":
            want = gen(p["root"], n, dtype, seed).float()
        else:
            per = p["per"]
            want = torch.cat([gen(r, n, dtype, seed).float()[rank * per:(rank + 1) * per] for r in allr])
        na = len(p.get("active", allr))
        tol = (0.05 * na ** 0.5 + 0.02 * float(want.abs().max())) if dtype == torch.bfloat16 else 1e-3
        ops.append(dict(kind=kind, p=p, src=src, x=x, want=want.to(dev), tol=tol,
                        out=torch.empty_like(x) if kind == "alltoall" else None))
    worst = torch.zeros(len(ops), device=dev)

    def run_all():
        for i, o in enumerate(ops):
            p, x = o["p"], o["x"]
            x.copy_(o["src"])
            res = x
            if o["kind"] == "allreduce":
                comm.all_reduce(x, op="sum", algo=p["algo"], active=p["active"])
            elif o["kind"] == "tree":
                comm.tree_collective(ALLREDUCE, x, op="sum", chunk_bytes=p["chunk"], active=p["active"])
            elif o["kind"] == "reduce":
                comm.reduce(x, root=p["root"], op="sum")
            elif o["kind"] == "broadcast":
                comm.broadcast(x, root=p["root"])
            else:
                res = comm.all_to_all(x, out=o["out"])
            worst[i] = torch.maximum(worst[i], (res.float() - o["want"]).abs().max() / o["tol"])

    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        run_all()                                   # eager pass first (allocations, first-touch)
        side.synchronize()
        comm.check()
        dist.barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            run_all()
        replays = max(1, (a.ops + len(ops) - 1) // len(ops))
        t0 = time.time()
        for _ in range(replays):
            g.replay()
        side.synchronize()
        dt = time.time() - t0
    comm.check()                                    # a timed-out wait latches the error word
    bad = [(i, float(w)) for i, w in enumerate(worst.cpu()) if not (w <= 1.0)]
    t = torch.tensor([len(bad)], device=dev)
    dist.all_reduce(t)
    for i, w in bad[:10]:
        print(f"[soak] rank {rank} op {i} {ops[i]['kind']} {ops[i]['p']}: error {w:.3g} x tolerance", flush=True)
    if rank == 0:
        print(f"[soak] world {world}: {len(ops)} ops/graph x {replays} replays = {len(ops) * replays} collectives in "
              f"{dt:.2f} s ({dt / (len(ops) * replays) * 1e6:.1f} us/op incl. re-init + verification), "
              f"failures: {int(t.item())}", flush=True)
    dist.barrier()
    comm.close()
    dist.destroy_process_group()
    sys.exit(1 if t.item() else 0)


if __name__ == "__main__":
    main()
