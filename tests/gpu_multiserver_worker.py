"""Multi-server data plane on one box: the ip table declares two 'servers' (first half / second half
of the ranks), so every collective takes the hierarchical path — our kernels inside each NVLink
domain, torch.distributed (NCCL) between the local roots. Launched by torchrun with >= 4 ranks."""
import os
import sys
from types import SimpleNamespace

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adapcc_b200 import ALLREDUCE  # noqa: E402
from adapcc_b200 import topology as topo  # noqa: E402
from adapcc_b200.adapcc import AdapCC  # noqa: E402


def main():
    os.environ.setdefault("ADAPCC_SHARED_FS", "1")      # the two 'servers' are halves of one box
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    work = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", f"multiserver_{world}")
    os.makedirs(os.path.join(work, "topology"), exist_ok=True)
    half = world // 2
    if rank == 0:
        topo.write_ip_table(os.path.join(work, "topology", "ip_table.txt"), ["10.0.0.1"] * half + ["10.0.0.2"] * (world - half))
    dist.barrier()
    args = SimpleNamespace(port=5000, strategy_file=os.path.join(work, "s.xml"), logical_graph=os.path.join(work, "lg.xml"),
                           entry_point=int(os.environ.get("ENTRY_POINT", 7)), parallel_degree=2, profile_freq=0, work_dir=work,
                           relay_control=False,
                           staging_mb=64)
    AdapCC.init(args, local, rank, world)
    AdapCC.setup(ALLREDUCE)
    comm = AdapCC.communicator
    assert not comm.single_server and len(comm.node_ranks) in (half, world - half)
    if rank == 0 and os.path.exists(args.strategy_file):
        print(open(args.strategy_file).read()[:900], flush=True)
        print(open(os.path.join(work, "topology", "topo_profile_0")).read(), flush=True)
    ok = True
    base = [torch.Generator().manual_seed(100 + r) for r in range(world)]
    data = torch.stack([torch.randn(100_003, generator=g) for g in base])
    for op, want in (("sum", data.sum(0)), ("avg", data.mean(0)), ("max", data.max(0).values)):
        t = data[rank].clone().to(dev)
        comm.all_reduce(t, op=op)
        comm.synchronize()
        ok &= bool(torch.allclose(t.cpu(), want, atol=1e-4))
    act = [0, world - 1]                                    # one active rank per server
    t = data[rank].clone().to(dev)
    comm.all_reduce(t, None, None, act, op="sum")
    comm.synchronize()
    want = data[act].sum(0) if rank in act else data[rank]
    ok &= bool(torch.allclose(t.cpu(), want, atol=1e-4))
    root = world - 1                                         # a non-local-root rank on the second server
    t = data[rank].clone().to(dev)
    comm.boardcast(t, root=root)
    comm.synchronize()
    ok &= bool(torch.equal(t.cpu(), data[root]))
    t = data[rank].clone().to(dev)
    comm.reduce(t, root=root)
    comm.synchronize()
    if rank == root:
        ok &= bool(torch.allclose(t.cpu(), data.sum(0), atol=1e-4))
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    print(f"[rank {rank}] multiserver {'OK' if ok else 'FAILED'}", flush=True)
    AdapCC.clear(ALLREDUCE)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
