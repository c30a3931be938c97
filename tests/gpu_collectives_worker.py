"""Multi-rank worker for the native collective kernels (launched by torchrun or by
tests/test_gpu_collectives.py). Checks every kernel variant against a plain fp32 PyTorch
reference computed from all ranks' inputs (gathered with NCCL/gloo or regenerated from the
shared seed), then optionally prints a small timing sweep.

Usage: torchrun --nproc-per-node N tests/gpu_collectives_worker.py [--sweep] [--quick]
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adapcc_b200.constants import ALLREDUCE, BOARDCAST, REDUCE  # noqa: E402
from adapcc_b200.runtime.native import NativeComm  # noqa: E402
from adapcc_b200.runtime.rendezvous import unique_name  # noqa: E402


def gen(rank, n, dtype, seed):
    g = torch.Generator(device="cpu").manual_seed(seed * 1000 + rank)
    return (torch.randn(n, generator=g, dtype=torch.float32) * 2).to(dtype)


def ref_reduce(world, n, dtype, seed, op, active, wire=None):
    xs = [gen(r, n, dtype, seed) for r in active]
    if wire is not None:
        xs = [x.to(wire) for x in xs]
    acc = torch.stack([x.float() for x in xs])
    if op == "max":
        return acc.max(0).values
    s = acc.sum(0)
    return s / len(active) if op == "avg" else s


def tol(dtype, wire, n_active):
    if (wire or dtype) in (torch.bfloat16,):
        return 4e-2 * max(1, n_active) ** 0.5, 2e-2
    if (wire or dtype) in (torch.float16,):
        return 1e-2 * max(1, n_active) ** 0.5, 4e-3
    return 1e-4, 1e-5


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default="")
    ap.add_argument("--sweep_max_log2", type=int, default=0, help="largest sweep message (log2 bytes); 0: 26 quick / 28 full")
    ap.add_argument("--sweep_step", type=int, default=0, help="log2 step between sweep sizes; 0: 2 quick / 1 full")
    ap.add_argument("--no_blocks_sweep", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    name = unique_name("worker")
    comm = NativeComm(name, rank, world, local, staging_bytes=256 << 20, heap_bytes=512 << 20)
    if rank == 0:
        print(f"[worker] world={world} symm={comm.symm_backend} multicast={comm.multicast} "
              f"heap_mc={comm.heap_multicast}", flush=True)
    failures = []
    n_checks = 0

    def check(tag, got, want, dtype, wire, na):
        nonlocal n_checks
        n_checks += 1
        atol, rtol = tol(dtype, wire, na)
        got = got.float().cpu()
        ok = torch.allclose(got, want, atol=atol, rtol=rtol)
        if not ok:
            err = (got - want).abs().max().item()
            failures.append(f"rank {rank} {tag}: max err {err:.4g}")
            print(f"[FAIL] rank {rank} {tag}: max err {err:.4g}", flush=True)

    sizes = [1, 7, 1024, 4097, 65536 + 3, 1 << 20, (1 << 22) + 5]
    if args.quick:
        sizes = [7, 4097, (1 << 20) + 3]
    algos = ["one_shot", "two_shot"] + (["nvls"] if comm.multicast else [])
    all_ranks = list(range(world))
    seed = 0
    # ---- direct allreduce: algos x dtypes x ops x sizes ------------------------------
    for dtype, wire in [(torch.float32, None), (torch.float32, "bfloat16"), (torch.bfloat16, None),
                        (torch.float16, None)]:
        wire_t = getattr(torch, wire) if wire else None
        for op in ["sum", "avg", "max"]:
            for algo in algos:
                if algo == "nvls" and op == "max" and (wire_t or dtype) == torch.float32:
                    continue
                for n in sizes:
                    seed += 1
                    x = gen(rank, n, dtype, seed).to(dev)
                    comm.all_reduce(x, op=op, algo=algo, wire=wire)
                    comm.check()
                    want = ref_reduce(world, n, dtype, seed, op, all_ranks, wire_t)
                    check(f"allreduce {algo} {dtype} wire={wire} {op} n={n}", x, want, dtype, wire_t, world)
    # ---- reduce-scatter = reduce with root = self, all-gather = one direct broadcast per shard -------------
    if True:
        for dtype in (torch.float32, torch.bfloat16):
            for algo in algos:
                if algo == "one_shot":
                    continue
                for n in (4097, (1 << 20) + 3):
                    seed += 1
                    for zc in (False, True):
                        if zc:
                            comm.heap_reset()
                            x = comm.symm_empty(n, dtype)
                            x.copy_(gen(rank, n, dtype, seed).to(dev))
                            dist.barrier()
                        else:
                            x = gen(rank, n, dtype, seed).to(dev)
                        lo, hi = comm.reduce_scatter_(x, op="sum", algo=algo)
                        comm.check()
                        want = ref_reduce(world, n, dtype, seed, "sum", all_ranks)
                        check(f"reduce_scatter {algo} {dtype} n={n} zc={zc} shard=[{lo},{hi})", x[lo:hi], want[lo:hi],
                              dtype, None, world)
                        comm.all_gather_(x)                 # reduce-scatter + all-gather == all-reduce
                        comm.check()
                        check(f"all_gather after reduce_scatter {algo} {dtype} n={n} zc={zc}", x, want, dtype, None,
                              world)
    # ---- low-latency path (flag-in-data; default on, ADAPCC_LL=0 turns the buffer off) --------------------
    if comm.has_ll:
        for dtype in (torch.float32, torch.bfloat16, torch.float16):
            for op in ("sum", "avg", "max"):
                for n in (1, 2, 3, 255, 1024, 4097, 8192):           # <= 32 KB, odd tails included
                    seed += 1
                    x = gen(rank, n, dtype, seed).to(dev)
                    comm.all_reduce(x, op=op, algo="ll")
                    comm.check()
                    check(f"allreduce ll {dtype} {op} n={n}", x, ref_reduce(world, n, dtype, seed, op, all_ranks),
                          dtype, None, world)
        # back-to-back ops without a host sync in between exercise the double-buffered slots; interleave a
        # barrier-protocol op so both counters move
        xs = [gen(rank, 1000 + i, torch.float32, 9000 + i).to(dev) for i in range(12)]
        for i, x in enumerate(xs):
            comm.all_reduce(x, op="sum", algo="ll" if i % 3 else "two_shot")
        comm.check()
        for i, x in enumerate(xs):
            check(f"allreduce ll back-to-back {i}", x, ref_reduce(world, 1000 + i, torch.float32, 9000 + i, "sum",
                                                                   all_ranks), torch.float32, None, world)
    # out-of-place + auto
    for n in sizes:
        seed += 1
        x = gen(rank, n, torch.float32, seed).to(dev)
        y = torch.empty_like(x)
        comm.all_reduce(x, out=y, op="sum", algo="auto")
        comm.check()
        check(f"allreduce auto oop n={n}", y, ref_reduce(world, n, torch.float32, seed, "sum", all_ranks),
              torch.float32, None, world)
        check(f"allreduce auto oop input intact n={n}", x, gen(rank, n, torch.float32, seed), torch.float32, None, 1)
    # ---- pipelined staged kernel (stager CTAs + link CTAs), forced on mid-size tensors --------------
    comm.set_tunable("pipe_min_bytes", 1 << 20)
    comm.set_tunable("pipe_nvls", 1)
    comm.set_tunable("pipe_piece_bytes", 1 << 19)
    for dtype, wire in [(torch.float32, None), (torch.float32, "bfloat16"), (torch.bfloat16, None)]:
        wire_t = getattr(torch, wire) if wire else None
        for algo in [a for a in algos if a != "one_shot"]:
            for n in [(1 << 20) + 3, (5 << 20) + 1]:
                seed += 1
                x = gen(rank, n, dtype, seed).to(dev)
                comm.all_reduce(x, op="avg", algo=algo, wire=wire)
                comm.check()
                check(f"pipelined {algo} {dtype} wire={wire} n={n}", x,
                      ref_reduce(world, n, dtype, seed, "avg", all_ranks, wire_t), dtype, wire_t, world)
    comm.set_tunable("pipe_min_bytes", 32 << 20)
    comm.set_tunable("pipe_piece_bytes", 16 << 20)
    # ---- zero-copy (symmetric heap) ----------------------------------------------------
    for dtype in [torch.float32, torch.bfloat16]:
        for algo in [a for a in algos if a != "one_shot"] + ["auto"]:
            for n in [4096, (1 << 20) + 8]:
                seed += 1
                comm.heap_reset()
                t = comm.symm_empty(n, dtype)
                t.copy_(gen(rank, n, dtype, seed).to(dev))
                torch.cuda.synchronize()
                dist.barrier()
                comm.all_reduce(t, op="avg", algo=algo)
                comm.check()
                check(f"zero-copy allreduce {algo} {dtype} n={n}", t,
                      ref_reduce(world, n, dtype, seed, "avg", all_ranks), dtype, None, world)
    # ---- zero-copy on a tensor that is NOT a whole number of 16-byte packs (a slice of a heap tensor): the partial
    # last pack must be stored element-wise — the bytes right after the slice belong to somebody else ----------------
    for dtype in [torch.float32, torch.bfloat16]:
        for algo in [a for a in algos if a != "one_shot"]:
            for n in [5, 4099, (1 << 18) + 3]:
                seed += 1
                comm.heap_reset()
                whole = comm.symm_empty(n + 64, dtype)
                whole.fill_(-7.0)
                t = whole[:n]
                t.copy_(gen(rank, n, dtype, seed).to(dev))
                torch.cuda.synchronize()
                dist.barrier()
                comm.all_reduce(t, op="sum", algo=algo)
                comm.check()
                check(f"zero-copy tail allreduce {algo} {dtype} n={n}", t,
                      ref_reduce(world, n, dtype, seed, "sum", all_ranks), dtype, None, world)
                check(f"zero-copy tail guard {algo} {dtype} n={n}", whole[n:], torch.full((64,), -7.0), torch.float32, None, 1)
                dist.barrier()
        seed += 1
        n = 4099
        comm.heap_reset()
        whole = comm.symm_empty(n + 64, dtype)
        whole.fill_(-7.0)
        t = whole[:n]
        t.copy_(gen(rank, n, dtype, seed).to(dev))
        torch.cuda.synchronize()
        dist.barrier()
        comm.broadcast(t, root=world - 1)
        comm.check()
        check(f"zero-copy tail broadcast {dtype}", t, gen(world - 1, n, dtype, seed).float(), dtype, None, 1)
        check(f"zero-copy tail broadcast guard {dtype}", whole[n:], torch.full((64,), -7.0), torch.float32, None, 1)
        dist.barrier()
    # ---- op sequence numbers stay in step when an op is split into staging-window pieces on the ranks that take
    # part while a non-participant runs one skip (small window, partial active set, then a tree op) -------------------
    if world >= 3:
        small = NativeComm(unique_name("small"), rank, world, local, staging_bytes=1 << 20, heap_bytes=0)
        small.load_strategy("<trees>" + "".join(
            "<root id='%d' ip='h'>%s</root>" % (o[0], "".join("<gpu id='%d' ip='h'>" % r for r in o[1:]) + "</gpu>" * (world - 1))
            for o in (list(range(world)), list(reversed(range(world))))) + "</trees>")
        act = [0, world - 1]
        for rep in range(2):
            seed += 1
            n = (5 << 18) + 11                                   # 5.2 MB of fp32: 6 pieces through a 1 MB window
            x = gen(rank, n, torch.float32, seed).to(dev)
            small.all_reduce(x, op="sum", algo="two_shot", active=act)
            small.check()
            want = ref_reduce(world, n, torch.float32, seed, "sum", act) if rank in act else gen(rank, n, torch.float32, seed)
            check(f"pieces + partial active set (rep {rep})", x, want, torch.float32, None, 2)
            seed += 1
            y = gen(rank, 70001, torch.float32, seed).to(dev)
            small.tree_collective(ALLREDUCE, y, op="sum", chunk_bytes=4096)
            small.check()
            check(f"tree op after split op (rep {rep})", y, ref_reduce(world, 70001, torch.float32, seed, "sum", all_ranks),
                  torch.float32, None, world * 2)
            seed += 1
            z = gen(rank, (3 << 18) + 5, torch.float32, seed).to(dev)    # tree op itself split into pieces
            small.tree_collective(ALLREDUCE, z, op="sum", chunk_bytes=65536)
            small.check()
            check(f"tree op in pieces (rep {rep})", z, ref_reduce(world, (3 << 18) + 5, torch.float32, seed, "sum", all_ranks),
                  torch.float32, None, world * 2)
        dist.barrier()
        small.close()
    # ---- strategy trees on heap tensors: reduced / broadcast in place (no staging pass) ---------------------------
    comm.load_strategy("<trees>" + "".join(
        "<root id='%d' ip='h'>%s</root>" % (o[0], "".join("<gpu id='%d' ip='h'>" % r for r in o[1:]) + "</gpu>" * (world - 1))
        for o in (list(range(world)), list(reversed(range(world))))) + "</trees>")
    for dtype in (torch.float32, torch.bfloat16):
        for n, chunk in [(4096, 1024), ((1 << 20) + 8, 1 << 16)]:
            seed += 1
            comm.heap_reset()
            t = comm.symm_empty(n, dtype)
            t.copy_(gen(rank, n, dtype, seed).to(dev))
            torch.cuda.synchronize()
            dist.barrier()
            comm.tree_collective(ALLREDUCE, t, op="sum", chunk_bytes=chunk)
            comm.check()
            check(f"tree in-place allreduce {dtype} n={n}", t, ref_reduce(world, n, dtype, seed, "sum", all_ranks), dtype,
                  dtype if dtype != torch.float32 else None, world * 2)
            dist.barrier()
    # ---- reduce-to-root and broadcast (direct) ---------------------------------------
    for root in sorted({0, world - 1}):
        for algo in algos:
            for n in [5, 70001]:
                seed += 1
                x = gen(rank, n, torch.float32, seed).to(dev)
                comm.reduce(x, root=root, op="sum", algo=algo)
                comm.check()
                if rank == root:
                    check(f"reduce {algo} root={root} n={n}", x,
                          ref_reduce(world, n, torch.float32, seed, "sum", all_ranks), torch.float32, None, world)
        for dtype in [torch.float32, torch.bfloat16]:
            for n in [3, 70001]:
                seed += 1
                x = gen(rank, n, dtype, seed).to(dev)
                comm.broadcast(x, root=root)
                comm.check()
                check(f"broadcast root={root} {dtype} n={n}", x, gen(root, n, dtype, seed).float(), dtype, None, 1)
    # ---- active subsets ----------------------------------------------------------------
    if world >= 3:
        subsets = [[0, world - 1], list(range(1, world))]
        for act in subsets:
            for algo in ["one_shot", "two_shot"]:
                seed += 1
                n = 100003
                x = gen(rank, n, torch.float32, seed).to(dev)
                comm.all_reduce(x, op="avg", algo=algo, active=act)
                comm.check()
                want = ref_reduce(world, n, torch.float32, seed, "avg", act) if rank in act else gen(rank, n, torch.float32, seed)
                check(f"subset {act} {algo}", x, want, torch.float32, None, len(act))
    # ---- strategy trees ----------------------------------------------------------------
    def chain_xml(order):
        s = ""
        for r in reversed(order[1:]):
            s = f"<gpu id='{r}' ip='h'>{s}</gpu>"
        return f"<root id='{order[0]}' ip='h'>{s}</root>"

    def bin_xml(order):
        def rec(i):
            kids = "".join(rec(c) for c in (2 * i + 1, 2 * i + 2) if c < len(order))
            tag = "root" if i == 0 else "gpu"
            return f"<{tag} id='{order[i]}' ip='h'>{kids}</{tag}>"
        return rec(0)

    orders = [list(range(world)), list(reversed(range(world)))]
    if world >= 4:
        orders.append([(r * 3 + 1) % world for r in range(world)] if world % 3 else orders[0][1:] + orders[0][:1])
    strategies = {
        "chain2": "<trees>" + "".join(chain_xml(o) for o in orders[:2]) + "</trees>",
        "binary": "<trees>" + "".join(bin_xml(o) for o in orders) + "</trees>",
    }
    if True:
        # random spanning trees (arbitrary fan-out and depth), the shapes the CPU property test covers
        import random as _random

        _rng = _random.Random(99)

        def random_tree_xml():
            order = list(range(world))
            _rng.shuffle(order)
            kids = {order[0]: []}
            for r in order[1:]:
                kids.setdefault(_rng.choice(list(kids)), []).append(r)
                kids.setdefault(r, [])

            def rec(a, tag):
                return f"<{tag} id='{a}' ip='h'>" + "".join(rec(c, "gpu") for c in kids[a]) + f"</{tag}>"
            return rec(order[0], "root")

        for k in range(6):
            xml = "<trees>" + "".join(random_tree_xml() for _ in range(1 + k % 3)) + "</trees>"
            comm.load_strategy(xml)
            for n, chunk in [(4097, 256), ((1 << 20) + 3, 1 << 16)]:
                seed += 1
                x = gen(rank, n, torch.float32, seed).to(dev)
                comm.tree_collective(ALLREDUCE, x, op="sum", chunk_bytes=chunk)
                comm.check()
                check(f"tree[random{k}] allreduce n={n}", x, ref_reduce(world, n, torch.float32, seed, "sum", all_ranks),
                      torch.float32, None, world * 2)
    for sname, xml in strategies.items():
        ntrees = comm.load_strategy(xml)
        for dtype, wire in [(torch.float32, None), (torch.float32, "bfloat16"), (torch.bfloat16, None)]:
            wire_t = getattr(torch, wire) if wire else None
            for n, chunk in [(16, 16), (70001, 4096), ((1 << 21) + 9, 1 << 18)]:
                seed += 1
                x = gen(rank, n, dtype, seed).to(dev)
                comm.tree_collective(ALLREDUCE, x, op="sum", wire=wire, chunk_bytes=chunk)
                comm.check()
                # tree sums round partial results to the wire dtype at every hop
                check(f"tree[{sname}] allreduce {dtype} wire={wire} n={n}", x,
                      ref_reduce(world, n, dtype, seed, "sum", all_ranks, wire_t), dtype, wire_t or dtype, world * 2)
        seed += 1
        n = 50001
        x = gen(rank, n, torch.float32, seed).to(dev)
        comm.tree_collective(BOARDCAST, x, chunk_bytes=8192)
        comm.check()
        # tree t broadcasts slice t from ITS root
        # (reference semantics: each tree owns one slice of the tensor)
        import math
        xml_roots = []
        import re
        for m in re.finditer(r"<root id='(\d+)'", xml):
            xml_roots.append(int(m.group(1)))
        epp = 4
        npacks = math.ceil(n / epp)
        per = math.ceil(npacks / ntrees)
        want = torch.empty(n)
        for t, r in enumerate(xml_roots):
            lo, hi = min(t * per, npacks) * epp, min(min((t + 1) * per, npacks) * epp, n)
            want[lo:hi] = gen(r, n, torch.float32, seed)[lo:hi]
        check(f"tree[{sname}] boardcast n={n}", x, want, torch.float32, None, 1)
        seed += 1
        x = gen(rank, n, torch.float32, seed).to(dev)
        comm.tree_collective(REDUCE, x, op="sum", chunk_bytes=8192)
        comm.check()
        full = ref_reduce(world, n, torch.float32, seed, "sum", all_ranks)
        mine = gen(rank, n, torch.float32, seed)
        for t, r in enumerate(xml_roots):
            lo, hi = min(t * per, npacks) * epp, min(min((t + 1) * per, npacks) * epp, n)
            if r == rank:
                mine[lo:hi] = full[lo:hi]
        check(f"tree[{sname}] reduce n={n}", x, mine, torch.float32, None, world)
        # relay control: forward and bypass modes over an active subset
        if world >= 3:
            act = [0, world - 1]
            for mode in (0, 1):
                comm.set_tunable("relay_mode", mode)
                seed += 1
                x = gen(rank, n, torch.float32, seed).to(dev)
                comm.tree_collective(ALLREDUCE, x, op="avg", chunk_bytes=8192, active=act)
                comm.check()
                want = ref_reduce(world, n, torch.float32, seed, "avg", act) if rank in act else gen(rank, n, torch.float32, seed)
                check(f"tree[{sname}] relay mode={mode} active={act}", x, want, torch.float32, None, 2)
            comm.set_tunable("relay_mode", 0)

    # ---- all-to-all (equal splits) ---------------------------------------------------------------
    for dtype in (torch.float32, torch.bfloat16):
        for per in (1, 37, 4096 + 3):
            blocks = torch.stack([torch.full((per,), float(rank * 16 + p)) for p in range(world)]).to(dtype).to(dev)
            out = comm.all_to_all(blocks.reshape(-1))
            comm.check()
            want = torch.stack([torch.full((per,), float(src * 16 + rank)) for src in range(world)]).reshape(-1)
            check(f"all_to_all {dtype} per={per}", out, want, dtype, None, 1)
            check(f"all_to_all input intact {dtype} per={per}", blocks.reshape(-1),
                  torch.stack([torch.full((per,), float(rank * 16 + p)) for p in range(world)]).reshape(-1), dtype, None, 1)

    # ---- persistent relay kernel: one launch forwards every bucket of a step ------------------
    if world >= 3:
        comm.load_strategy(strategies["chain2"])
        comm.set_tunable("relay_mode", 0)
        act = [0, world - 1]
        sizes_b = [40001, 262144 + 7, 1000]
        chunks_b = [4096, 65536, 256]
        seed += 1
        xs = [gen(rank, n, torch.float32, seed * 10 + i).to(dev) for i, n in enumerate(sizes_b)]
        if rank in act:
            for x, cb in zip(xs, chunks_b):
                comm.tree_collective(ALLREDUCE, x, op="avg", chunk_bytes=cb, active=act)
        else:
            comm.tree_relay_persistent(sizes_b, chunks_b, wire="float32", op="avg", active=act)
        comm.check()
        for i, (x, n) in enumerate(zip(xs, sizes_b)):
            want = ref_reduce(world, n, torch.float32, seed * 10 + i, "avg", act) if rank in act else gen(rank, n, torch.float32, seed * 10 + i)
            check(f"persistent relay bucket {i}", x, want, torch.float32, None, 2)
        # and the op sequence stayed in step: a normal collective still works afterwards
        seed += 1
        x = gen(rank, 5000, torch.float32, seed).to(dev)
        comm.tree_collective(ALLREDUCE, x, op="sum", chunk_bytes=4096)
        comm.check()
        check("tree after persistent relay", x, ref_reduce(world, 5000, torch.float32, seed, "sum", all_ranks), torch.float32, None, world)

    # ---- expert-parallel MoE exchange (dispatch / combine over peer memory) ---------------------
    if world >= 2:
        from adapcc_b200.models.moe import MoEMLP
        from adapcc_b200.parallel.expert_parallel import ExpertExchange

        E_local, d, h, T, k = 2, 64, 128, 64, 2
        torch.manual_seed(77)                                   # identical full expert set everywhere
        full = MoEMLP(E_local * world, d, h, top_k=k).to(dev).bfloat16()
        comm.heap_reset()
        ex = ExpertExchange(comm, E_local, full.capacity(T) * 2, d)
        moe = MoEMLP(E_local, d, h, top_k=k, world_size=world, exchange=ex).to(dev).bfloat16()
        with torch.no_grad():
            moe.gate.load_state_dict(full.gate.state_dict())
            sl = slice(rank * E_local, (rank + 1) * E_local)
            for name in ("w1", "b1", "w2", "b2"):
                getattr(moe, name).copy_(getattr(full, name)[sl])
        xt = gen(rank, T * d, torch.float32, 4242).view(T, d).to(dev).bfloat16()
        x1, x2 = xt.clone().requires_grad_(True), xt.clone().requires_grad_(True)
        y_ref = full(x1)                                        # all experts local: the oracle
        y = moe(x2)
        comm.check()
        check("moe expert-parallel forward", y.flatten(), y_ref.detach().float().flatten().cpu(), torch.bfloat16, None, 4)
        g = gen(rank, T * d, torch.float32, 4243).view(T, d).to(dev).bfloat16()
        y_ref.backward(g)
        y.backward(g)
        comm.check()
        check("moe expert-parallel grad_x", x2.grad.flatten(), x1.grad.detach().float().flatten().cpu(), torch.bfloat16, None, 4)

    torch.cuda.synchronize()
    dist.barrier()
    fl = [None] * world
    dist.all_gather_object(fl, failures)
    all_fail = [f for sub in fl for f in sub]
    if rank == 0:
        print(f"[worker] checks per rank: {n_checks}; failures: {len(all_fail)}", flush=True)
        for f in all_fail[:40]:
            print("   ", f)

    # ---- timing sweep ------------------------------------------------------------------
    results = []
    if args.sweep:
        comm.load_strategy(strategies["binary"])
        top = args.sweep_max_log2 or (26 if args.quick else 28)
        sweep_sizes = [1 << p for p in range(10, top + 1, args.sweep_step or (2 if args.quick else 1))]  # bytes
        side = torch.cuda.Stream()

        def timeit(fn, iters, graph=True):
            """Device time per call, max over ranks. graph=True replays a captured CUDA graph of
            `iters` back-to-back calls, so Python/launch overhead is excluded for both NCCL and us."""
            with torch.cuda.stream(side):
                for _ in range(3):
                    fn()
                side.synchronize()
                g = None
                if graph:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        for _ in range(iters):
                            fn()
                    g.replay()
                    side.synchronize()
                best = 1e9
                for _ in range(3):
                    dist.barrier()
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record(side)
                    if g is not None:
                        g.replay()
                    else:
                        for _ in range(iters):
                            fn()
                    e.record(side)
                    side.synchronize()
                    best = min(best, s.elapsed_time(e) / iters)
            t = torch.tensor([best], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item() * 1e-3

        for nbytes in sweep_sizes:
            n = nbytes // 4
            x = torch.randn(n, device=dev)
            comm.heap_reset()
            hz = comm.symm_empty(n, torch.float32) if nbytes <= comm.heap_bytes else None
            iters = 40 if nbytes <= (1 << 22) else (10 if nbytes <= (1 << 26) else 4)
            row = {"bytes": nbytes}
            row["nccl"] = timeit(lambda: dist.all_reduce(x), iters)
            row["nccl_eager"] = timeit(lambda: dist.all_reduce(x), iters, graph=False)
            for algo in algos + ["auto"]:
                if algo == "one_shot" and nbytes > (4 << 20):
                    continue
                row[algo] = timeit(lambda: comm.all_reduce(x, algo=algo), iters)
                if hz is not None and algo not in ("one_shot",):
                    row[algo + "_zc"] = timeit(lambda: comm.all_reduce(hz, algo=algo), iters)
            if comm.has_ll and nbytes <= 32768:
                row["ll"] = timeit(lambda: comm.all_reduce(x, algo="ll"), iters)
            row["auto_eager"] = timeit(lambda: comm.all_reduce(x, algo="auto"), iters, graph=False)
            row["bf16wire_auto"] = timeit(lambda: comm.all_reduce(x, algo="auto", wire="bfloat16"), iters)
            if nbytes >= (1 << 16):
                row["tree"] = timeit(lambda: comm.tree_collective(ALLREDUCE, x, chunk_bytes=4 << 20), iters)
                if hz is not None:
                    row["tree_zc"] = timeit(lambda: comm.tree_collective(ALLREDUCE, hz, chunk_bytes=4 << 20), iters)
            comm.check()
            results.append(row)
            if rank == 0:
                f = 2 * (world - 1) / world
                print("[sweep] %10d B " % nbytes + " ".join(
                    f"{k}={v * 1e6:7.1f}us({nbytes * f / v / 1e9:6.1f})" for k, v in row.items() if k != "bytes"),
                    flush=True)
        # reduce / broadcast / all-to-all next to NCCL (nccl-tests' reduce, broadcast and alltoall,
        # /root/reference/nccl-perf/benchmark/src/{reduce,broadcast,alltoall}.cu; busbw factor 1, resp. (n-1)/n)
        for nbytes in [1 << p for p in range(10, min(top, 28) + 1, 3 if args.quick else 2)]:
            n = nbytes // 4
            x = torch.randn(n, device=dev)
            iters = 40 if nbytes <= (1 << 22) else (10 if nbytes <= (1 << 26) else 4)
            row = {"bytes": nbytes, "prims": 1}
            row["reduce_nccl"] = timeit(lambda: dist.reduce(x, dst=0), iters)
            row["reduce"] = timeit(lambda: comm.reduce(x, root=0, op="sum"), iters)
            row["bcast_nccl"] = timeit(lambda: dist.broadcast(x, src=0), iters)
            row["bcast"] = timeit(lambda: comm.broadcast(x, root=0), iters)
            comm.heap_reset()
            if nbytes <= comm.heap_bytes:
                hz = comm.symm_empty(n, torch.float32)
                row["reduce_zc"] = timeit(lambda: comm.reduce(hz, root=0, op="sum"), iters)
                row["bcast_zc"] = timeit(lambda: comm.broadcast(hz, root=0), iters)
            if n % world == 0 and nbytes <= (128 << 20):
                y = torch.empty_like(x)
                row["a2a_nccl"] = timeit(lambda: dist.all_to_all_single(y, x), iters)
                row["a2a"] = timeit(lambda: comm.all_to_all(x, out=y), iters)
            comm.check()
            results.append(row)
            if rank == 0:
                fa = (world - 1) / world
                print("[prims] %10d B " % nbytes + " ".join(
                    f"{k}={v * 1e6:7.1f}us({nbytes * (fa if k.startswith('a2a') else 1.0) / v / 1e9:6.1f})"
                    for k, v in row.items() if k not in ("bytes", "prims")), flush=True)
        # CTA-count sensitivity at large sizes (zero-copy paths)
        for nbytes in ([] if args.no_blocks_sweep else [1 << 24, 1 << 26]):
            n = nbytes // 4
            comm.heap_reset()
            hz = comm.symm_empty(n, torch.float32)
            for blocks in [16, 32, 64, 96, 128, 148]:
                comm.set_tunable("max_blocks", blocks)
                comm.set_tunable("tree_blocks", blocks)
                row = {"bytes": nbytes, "blocks": blocks}
                for algo in [a for a in algos if a != "one_shot"]:
                    row[algo + "_zc"] = timeit(lambda: comm.all_reduce(hz, algo=algo), 8)
                x = torch.randn(n, device=dev)
                row["two_shot"] = timeit(lambda: comm.all_reduce(x, algo="two_shot"), 8)
                row["tree"] = timeit(lambda: comm.tree_collective(ALLREDUCE, x, chunk_bytes=4 << 20), 8)
                comm.check()
                results.append(row)
                if rank == 0:
                    f = 2 * (world - 1) / world
                    print(f"[blocks] {nbytes} B blocks={blocks} " + " ".join(
                        f"{k}={v * 1e6:7.1f}us({nbytes * f / v / 1e9:6.1f})" for k, v in row.items()
                        if k not in ("bytes", "blocks")), flush=True)
            comm.set_tunable("max_blocks", 64)
            comm.set_tunable("tree_blocks", 64)
    if rank == 0 and args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as fh:
            json.dump({"world": world, "symm": comm.symm_backend, "multicast": comm.multicast,
                       "checks": n_checks, "failures": all_fail, "sweep": results}, fh, indent=1)
    dist.barrier()
    comm.close()
    dist.destroy_process_group()
    sys.exit(1 if all_fail else 0)


if __name__ == "__main__":
    main()
