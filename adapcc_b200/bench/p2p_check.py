"""Symmetric-memory / P2P sanity check — the role of the reference's CUDA-aware-MPI ping-pong
(/root/reference/units-test/check-p2p/check_mpi_p2p.cu, check_mpi_connect.py): can every rank map
every peer's buffer, do peer stores become visible, does the multicast mapping reduce?

    torchrun --nproc-per-node 8 -m adapcc_b200.bench.p2p_check
"""
import os

import torch
import torch.distributed as dist

from ..runtime.native import NativeComm
from ..runtime.rendezvous import unique_name


def main():
    import argparse

    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--native", action="store_true",
                    help="run the stand-alone native binary adapcc_b200/_C/check_p2p instead (one process, all GPUs)")
    a, rest = ap.parse_known_args()
    if a.native:
        import subprocess

        from ..build import OUT_DIR, build
        build()
        raise SystemExit(subprocess.call([str(OUT_DIR / "check_p2p"), *rest]))
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    comm = NativeComm(unique_name("p2p"), rank, world, local, staging_bytes=8 << 20, heap_bytes=8 << 20)
    print(f"[rank {rank}] symmetric memory: backend={comm.symm_backend} multicast={comm.multicast}", flush=True)
    ok = True
    x = torch.full((1 << 20,), float(rank + 1), device=dev)
    comm.all_reduce(x, op="sum", algo="two_shot")           # peer loads + peer stores
    comm.check()
    ok &= bool((x == world * (world + 1) / 2).all())
    y = torch.full((4096,), float(rank), device=dev)
    comm.broadcast(y, root=world - 1)                        # multimem.st when multicast is bound
    comm.check()
    ok &= bool((y == world - 1).all())
    if comm.multicast and world > 1:
        z = torch.ones(1 << 18, device=dev)
        comm.all_reduce(z, algo="nvls")                      # multimem.ld_reduce through the switch
        comm.check()
        ok &= bool((z == world).all())
    print(f"[rank {rank}] {'P2P OK' if ok else 'P2P FAILED'}", flush=True)
    if world > 1:
        dist.barrier()
    comm.close()
    if world > 1:
        dist.destroy_process_group()
    raise SystemExit(0 if ok else 1)


if __name__ == "__main__":
    main()
