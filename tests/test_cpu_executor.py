"""BASELINE config 1: all_reduce correctness at world_size=2 on CPU/gloo with the 4-GPU strategy
file (plumbing, no GPU) — plus reduce / boardcast / relay subsets at world_size=4 through the public
API (AdapCC.init -> setup -> communicator.all_reduce)."""
import os
import sys
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import STRATEGY_4, STRATEGY_TEST

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, xml, tmp, result_q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from types import SimpleNamespace

        from adapcc_b200 import ALLREDUCE
        from adapcc_b200.adapcc import AdapCC

        sf = os.path.join(tmp, "strategy.xml")
        if rank == 0:
            with open(sf, "w") as f:
                f.write(xml)
        dist.barrier()
        args = SimpleNamespace(port=5000, strategy_file=sf, logical_graph=os.path.join(tmp, "lg.xml"),
                               entry_point=-1, parallel_degree=4, profile_freq=500, backend="gloo",
                               work_dir=tmp, relay_control=False)
        AdapCC.init(args, rank, rank, world)
        AdapCC.setup(ALLREDUCE)
        comm = AdapCC.communicator
        ok = True
        # the reference's primitive benchmark: ones(16)*i, chunk 8 bytes -> every rank sees world*i
        for i in (1, 2):
            t = torch.ones(16) * i
            out = comm.all_reduce(t, 16, 8, list(range(world)))
            ok &= bool(torch.equal(out, torch.full((16,), float(world * i))))
        # odd sizes / tails (the reference drops them)
        for n, chunk in [(1, 16), (1001, 64), (70001, 4096)]:
            g = torch.Generator().manual_seed(n)
            base = torch.randn(world, n, generator=g)
            t = base[rank].clone()
            comm.all_reduce(t, n, chunk, list(range(world)))
            ok &= bool(torch.allclose(t, base.sum(0), atol=1e-5))
            t = base[rank].clone()
            comm.all_reduce(t, n, chunk, list(range(world)), op="avg")
            ok &= bool(torch.allclose(t, base.mean(0), atol=1e-5))
        if world >= 4:
            n = 5003
            g = torch.Generator().manual_seed(7)
            base = torch.randn(world, n, generator=g)
            for mode in ("forward", "bypass"):
                comm.relay_mode = 0 if mode == "forward" else 1
                act = [0, 2]
                t = base[rank].clone()
                comm.all_reduce(t, n, 512, act)
                want = base[act].sum(0) if rank in act else base[rank]
                ok &= bool(torch.allclose(t, want, atol=1e-5))
            comm.relay_mode = 0
            # reduce: slice t of the result lands on tree t's root only
            from adapcc_b200.strategy import Strategy, slice_bounds

            s = Strategy.from_xml(xml, world)
            b = slice_bounds(n, len(s.trees), 4)
            t = base[rank].clone()
            comm.reduce(t, n, 512, list(range(world)))
            want = base[rank].clone()
            for ti, tr in enumerate(s.trees):
                if tr.root == rank:
                    want[b[ti]:b[ti + 1]] = base.sum(0)[b[ti]:b[ti + 1]]
            ok &= bool(torch.allclose(t, want, atol=1e-5))
            t = base[rank].clone()
            comm.boardcast(t, n, 512)
            want = torch.empty(n)
            for ti, tr in enumerate(s.trees):
                want[b[ti]:b[ti + 1]] = base[tr.root][b[ti]:b[ti + 1]]
            ok &= bool(torch.allclose(t, want))
        AdapCC.clear(ALLREDUCE)
        result_q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def _run(world, xml):
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as tmp:
        procs = [ctx.Process(target=_worker, args=(r, world, port, xml, tmp, q)) for r in range(world)]
        [p.start() for p in procs]
        [p.join(120) for p in procs]
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        res = dict(q.get(timeout=5) for _ in range(world))
    assert all(res.values()), res


def test_allreduce_world2_strategy4_gloo():
    _run(2, STRATEGY_4)


def test_collectives_and_relay_world4_gloo():
    _run(4, STRATEGY_TEST)
