"""Unusual inputs through the public API on CPU/gloo at an odd world size (3) with no strategy file: max, partial
size, float64 / bfloat16 / int64 tensors, empty tensor, non-contiguous view, default arguments, 2-D tensor."""
import os
import sys
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import free_port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fuzz_worker(rank, world, port, tmp, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from types import SimpleNamespace
    from adapcc_b200 import ALLREDUCE
    from adapcc_b200.adapcc import AdapCC
    args = SimpleNamespace(port=5000, strategy_file=os.path.join(tmp, "none.xml"), logical_graph=os.path.join(tmp, "lg.xml"),
                           entry_point=-1, parallel_degree=2, profile_freq=500, backend="gloo", work_dir=tmp, relay_control=False)
    AdapCC.init(args, rank, rank, world); AdapCC.setup(ALLREDUCE)
    comm = AdapCC.communicator
    problems = []
    def case(name, fn):
        try:
            fn()
        except Exception as e:
            problems.append(f"{name}: {type(e).__name__}: {e}")
    g = torch.Generator().manual_seed(1)
    base = torch.randn(world, 37, generator=g)
    def c_max():
        t = base[rank].clone(); comm.all_reduce(t, 37, 16, list(range(world)), op="max")
        assert torch.allclose(t, base.max(0).values), "max wrong"
    def c_partial():
        t = base[rank].clone(); comm.all_reduce(t, 10, 16, list(range(world)))
        assert torch.allclose(t[:10], base.sum(0)[:10]) and torch.equal(t[10:], base[rank][10:]), "partial size wrong"
    def c_f64():
        t = base[rank].double(); comm.all_reduce(t, 37, 16, list(range(world)))
        assert torch.allclose(t, base.double().sum(0)), "f64 wrong"
    def c_bf16():
        t = base[rank].bfloat16(); comm.all_reduce(t, 37, 16, list(range(world)))
        assert torch.allclose(t.float(), base.bfloat16().float().sum(0), atol=0.1), "bf16 wrong"
    def c_int():
        t = torch.arange(9) + rank; comm.all_reduce(t, 9, 16, list(range(world)))
        assert torch.equal(t, torch.arange(9) * world + sum(range(world))), "int wrong"
    def c_zero():
        t = torch.zeros(0); comm.all_reduce(t, 0, 16, list(range(world)))
    def c_noncontig():
        m = torch.stack([base[rank], base[rank]], 1)           # [37, 2]
        t = m[:, 0]                                            # stride 2
        comm.all_reduce(t, 37, 16, list(range(world)))
        assert torch.allclose(m[:, 0], base.sum(0)), "non-contiguous wrong"
    def c_none_args():
        t = base[rank].clone(); comm.all_reduce(t)
        assert torch.allclose(t, base.sum(0)), "defaults wrong"
    def c_2d():
        t = base[rank].clone().view(1, 37); AdapCC.allreduce(t)
        assert torch.allclose(t.view(-1), base.sum(0)), "2d wrong"
    for n, f in list(locals().items()):
        if n.startswith("c_"):
            case(n, f)
            dist.barrier()
    AdapCC.clear(ALLREDUCE)
    q.put((rank, problems))
    dist.destroy_process_group()



def test_public_api_handles_unusual_inputs_world3():
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    tmp = tempfile.mkdtemp()
    port = free_port()
    procs = [ctx.Process(target=_fuzz_worker, args=(r, world, port, tmp, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=20)
    assert sorted(r for r, _ in results) == list(range(world))
    assert all(not problems for _, problems in results), results
