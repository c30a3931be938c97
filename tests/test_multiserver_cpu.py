"""Two emulated servers (2 x 2 ranks, two addresses in the ip table) through the whole control plane on CPU/gloo:
per-server detection, a two-server logical graph, inter-server profiling, parallel trees rooted on different servers
(the reference's parity rule), collectives over all ranks and over one active rank per server."""
import os
import sys
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import free_port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ms_worker(rank, world, port, tmp, entry, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), ADAPCC_SHARED_FS="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from types import SimpleNamespace
    from adapcc_b200 import ALLREDUCE, topology as topo
    from adapcc_b200.adapcc import AdapCC
    from adapcc_b200.strategy import Strategy
    problems = []
    try:
        half = world // 2
        os.makedirs(os.path.join(tmp, "topology"), exist_ok=True)
        if rank == 0:
            topo.write_ip_table(os.path.join(tmp, "topology", "ip_table.txt"), ["10.0.0.1"] * half + ["10.0.0.2"] * (world - half))
        dist.barrier()
        args = SimpleNamespace(port=5000, strategy_file=os.path.join(tmp, "strategy", "ms.xml"), logical_graph=os.path.join(tmp, "topology", "lg.xml"),
                               entry_point=entry, parallel_degree=2, profile_freq=0, work_dir=tmp, relay_control=False, backend="gloo")
        AdapCC.init(args, rank % half, rank, world)
        AdapCC.setup(ALLREDUCE)
        comm = AdapCC.communicator
        if comm.single_server: problems.append("single_server is True")
        if rank == 0:
            s = Strategy.from_file(args.strategy_file); s.validate(world)
            if entry == 6 and open(args.logical_graph).read().count("<server ") != 2: problems.append("logical graph")
            if {t.root for t in s.trees} != {0, half}: problems.append(f"roots {[t.root for t in s.trees]}")
        g = [torch.Generator().manual_seed(100 + r) for r in range(world)]
        data = torch.stack([torch.randn(1003, generator=x) for x in g])
        for op, want in (("sum", data.sum(0)), ("avg", data.mean(0)), ("max", data.max(0).values)):
            t = data[rank].clone(); comm.all_reduce(t, op=op)
            if not torch.allclose(t, want, atol=1e-5): problems.append(f"allreduce {op}")
        act = [0, world - 1]
        t = data[rank].clone(); comm.all_reduce(t, None, None, act)
        want = data[act].sum(0) if rank in act else data[rank]
        if not torch.allclose(t, want, atol=1e-5): problems.append("subset")
        AdapCC.clear(ALLREDUCE)
    except Exception as e:
        import traceback; problems.append(traceback.format_exc()[-600:])
    q.put((rank, problems))
    dist.destroy_process_group()



@pytest.mark.parametrize("entry", [6, 7])
def test_two_server_workflow_on_cpu(entry):
    world = 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    tmp = tempfile.mkdtemp()
    port = free_port()
    procs = [ctx.Process(target=_ms_worker, args=(r, world, port, tmp, entry, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=20)
    assert sorted(r for r, _ in results) == list(range(world))
    assert all(not problems for _, problems in results), results
