"""Link-contention probe — the reference measures a 40 MB CUDA-aware MPI transfer while a persistent
400 MB background flow shares (or does not share) the NIC (/root/reference/nccl-perf/contention/
flow.cu, flow_dedicate.cu). On an NVSwitch box the question becomes: how much does a peer read lose
when another flow targets the same source GPU (shared egress port) versus a different one?

  foreground : rank 0 reads `--mb` MiB from rank 1 (kernel peer loads, device-timed)
  background : rank 2 continuously reads from rank 1 (shared) or from rank 3 (dedicated)

    torchrun --nproc-per-node 4 -m adapcc_b200.bench.contention
"""
import argparse
import os

import torch
import torch.distributed as dist

from ..runtime.native import NativeComm
from ..runtime.rendezvous import unique_name


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=256)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    if world < 4:
        raise SystemExit("needs 4 ranks")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    comm = NativeComm(unique_name("contention"), rank, world, local, staging_bytes=a.mb << 20)
    n = (a.mb << 20) // 4
    out = torch.empty(n, device=dev)

    def peer_tensor(r):
        from ..runtime.native import _CudaArray

        ptr = comm.lib.adapcc_ctx_peer_staging_ptr(comm.handle, r)
        return torch.as_tensor(_CudaArray(ptr, n * 4, comm), device=dev).view(torch.float32)

    results = {}
    for mode, bg_src in (("alone", None), ("shared_source", 1), ("dedicated_source", 3)):
        dist.barrier()
        stop = torch.zeros(1, device=dev)
        if rank == 0:
            src = peer_tensor(1)
            for _ in range(2):
                out.copy_(src)
            torch.cuda.synchronize()
        dist.barrier()
        if rank == 2 and bg_src is not None:
            src = peer_tensor(bg_src)
            for _ in range(40):                              # background flow
                out.copy_(src)
        if rank == 0:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                out.copy_(src)
            e1.record()
            torch.cuda.synchronize()
            results[mode] = n * 4 * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e9
        torch.cuda.synchronize()
        dist.barrier()
        _ = stop
    if rank == 0:
        for k, v in results.items():
            print(f"foreground 0<-1 {k:17s}: {v:7.1f} GB/s")
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
