"""Multi-GPU end-to-end workflow through the public API (launched by torchrun):
AdapCC.init(entry_point=6) = native topology detection + link profiling + synthesis, setup,
primitives, DDP + cuda_allreduce_hook with a forced straggler (relay control), zero-copy DDP buckets,
reconstruct_topology. Prints the profiled link matrix and relay statistics; exits non-zero on any
numerical mismatch."""
import os
import sys
import time
from types import SimpleNamespace

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adapcc_b200 import ALLREDUCE  # noqa: E402
from adapcc_b200.adapcc import AdapCC  # noqa: E402
from adapcc_b200.parallel.ddp import rebuild_buckets, wrap_ddp  # noqa: E402


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    work = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", f"workflow_{world}")
    os.makedirs(os.path.join(work, "strategy"), exist_ok=True)
    args = SimpleNamespace(port=5000, strategy_file=os.path.join(work, "strategy", "auto.xml"),
                           logical_graph=os.path.join(work, "topology", "logical_graph.xml"), entry_point=6,
                           parallel_degree=4, profile_freq=0, work_dir=work, relay_threshold=0.03, policy="auto",
                           heap_mb=256, staging_mb=64, relay_mode=os.environ.get("RELAY_MODE", "bypass"),
                           algo=os.environ.get("ALGO", "auto"), coordinator_port=50061)
    ok = True
    t0 = time.time()
    AdapCC.init(args, local, rank, world)
    AdapCC.setup(ALLREDUCE)
    comm = AdapCC.communicator
    if rank == 0:
        print(f"[workflow] detect+profile+synth+setup: {(time.time() - t0) * 1e3:.0f} ms", flush=True)
        print(open(os.path.join(work, "topology", "topo_detect_0.xml")).read()[:1500], flush=True)
        for r in range(min(world, 2)):
            print(open(os.path.join(work, "topology", f"topo_profile_{r}")).read(), flush=True)
        print(open(os.path.join(work, "topology", "tunables.json")).read(), flush=True)
        print(open(args.strategy_file).read()[:600], flush=True)
    # primitives through the reference-style calls
    for n in (16, 100_003, 4 << 20):
        t = torch.full((n,), float(rank + 1), device=dev)
        comm.all_reduce(t, n, None, list(range(world)))
        comm.synchronize()
        ok &= bool((t == world * (world + 1) / 2).all())
    t = torch.full((5000,), float(rank), device=dev)
    comm.boardcast(t, 5000)
    comm.synchronize()
    ok &= bool((t == 0).all())
    # reduce-scatter / all-gather through the public API (native reduce with root = self + per-shard broadcasts)
    n = 100_003
    t = torch.arange(n, dtype=torch.float32, device=dev) % 97 * (rank + 1)
    full = torch.arange(n, dtype=torch.float32, device=dev) % 97 * (world * (world + 1) / 2)
    lo, hi = comm.reduce_scatter(t)
    comm.synchronize()
    ok &= 0 <= lo < hi <= n and bool(torch.equal(t[lo:hi], full[lo:hi]))
    comm.all_gather(t)
    comm.synchronize()
    ok &= bool(torch.equal(t, full))
    # DDP + hook, buckets in the symmetric heap, a straggler from step 2 on
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(512, 2048), torch.nn.ReLU(), torch.nn.Linear(2048, 512)).to(dev)
    ddp = wrap_ddp(model, comm, local, bucket_cap_mb=2)
    opt = torch.optim.SGD(ddp.parameters(), lr=0.01)
    straggler = world - 1 if world > 2 else -1
    for step in range(6):
        comm.update_relay(step)
        loss = ddp(torch.randn(64, 512, device=dev)).pow(2).mean()
        opt.zero_grad(set_to_none=False)
        if rank == straggler and step >= 2:
            time.sleep(0.25)
        loss.backward()
        opt.step()
        if step == 0:
            rebuild_buckets(ddp, comm)
        torch.cuda.synchronize()
        print(f"[rank {rank}] step {step} loss {loss.item():.4f} active {comm.active_gpus}", flush=True)
    comm.synchronize()
    stats = comm.stats
    print(f"[rank {rank}] relay_steps={stats['relay_steps']} hook_rpc_ms="
          f"{[round(x * 1e3, 2) for x in stats['hook_rpc_s']]}", flush=True)
    if straggler >= 0:
        ok &= (stats["relay_steps"] > 0) == (rank == straggler)
    t0 = time.time()
    AdapCC.reconstruct_topology(args, ALLREDUCE)
    comm = AdapCC.communicator
    t = torch.ones(1000, device=dev)
    comm.all_reduce(t, 1000)
    comm.synchronize()
    ok &= bool((t == world).all())
    if rank == 0:
        print(f"[workflow] reconstruct_topology: {(time.time() - t0) * 1e3:.0f} ms", flush=True)
    # training continues with the SAME DDP object: its hook is bound to the cleared communicator (forwards to the live
    # one) and its gradient buckets live in the symmetric heap of the native context that survived the reconstruct
    for step in range(6, 9):
        comm.update_relay(step)
        loss = ddp(torch.randn(64, 512, device=dev)).pow(2).mean()
        opt.zero_grad(set_to_none=False)
        loss.backward()
        opt.step()
        torch.cuda.synchronize()
    comm.synchronize()
    if straggler < 0:                      # with relay steps the late rank's replica legitimately differs (BSP relays)
        w0 = model[0].weight.detach().clone()
        dist.broadcast(w0, src=0)
        ok &= bool(torch.allclose(w0, model[0].weight.detach(), atol=1e-5))
    ok &= len(comm.stats["hook_rpc_s"]) >= 1
    if rank == 0:
        print(f"[workflow] DDP across reconstruct: hook forwarded, {len(comm.stats['hook_rpc_s'])} negotiated steps", flush=True)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    AdapCC.clear(ALLREDUCE)
    if rank == 0:
        print("[workflow] " + ("OK" if flag.item() == 1 else "FAILED"), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
