"""``AdapCC.alltoall`` — dense all-to-all of equal splits. The reference lists ALLTOALL=4 among its
primitives and exposes ``AdapCC.alltoall`` but it forwards to a method that does not exist
(/root/reference/adapcc.py:59-61). GPU tensors use the native push kernel (in-kernel NVLink stores
into the peers' windows); CPU tensors fall back to ``torch.distributed.all_to_all_single`` when the
backend supports it, else to pairwise send/recv."""
from __future__ import annotations

import torch
import torch.distributed as dist


def all_to_all_single(communicator, tensor: torch.Tensor, size=None) -> torch.Tensor:
    flat = tensor.reshape(-1) if size is None else tensor.reshape(-1)[: int(size)]
    world = communicator.world_size
    if flat.numel() % world:
        raise ValueError(f"alltoall: {flat.numel()} elements do not split over {world} ranks")
    if flat.is_cuda:
        native = communicator._ensure_native()
        return native.all_to_all(flat.contiguous())
    out = torch.empty_like(flat)
    per = flat.numel() // world
    rank = communicator.world_rank
    out[rank * per:(rank + 1) * per] = flat[rank * per:(rank + 1) * per]
    reqs = []
    for p in range(world):
        if p == rank:
            continue
        reqs.append(dist.isend(flat[p * per:(p + 1) * per].contiguous(), dst=p))
    for p in range(world):
        if p == rank:
            continue
        buf = torch.empty(per, dtype=flat.dtype)
        dist.recv(buf, src=p)
        out[p * per:(p + 1) * per] = buf
    for r in reqs:
        r.wait()
    return out
