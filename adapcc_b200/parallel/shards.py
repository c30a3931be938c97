"""``AdapCC.reducescatter`` / ``AdapCC.allgather`` — the two primitive ids the reference declares but never
implements (ALLGATHER = 3, REDUCESCATTER = 5, /root/reference/commu.py:19-26). In place on one buffer, with the shard
layout of the direct kernels (``parallel.engine.shard_of``: 16-byte packs dealt out in ``world`` contiguous slices), so
``reducescatter`` followed by ``allgather`` is an all-reduce and the ZeRO-1 engine's buckets use the same partition.

GPU tensors: the native communicator (reduce with root = self over peer memory / NVLS; one multicast-store broadcast
per shard). CPU tensors (gloo): one ``dist.reduce`` / ``dist.broadcast`` per shard."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist

from .engine import shard_of


def _flat(tensor: torch.Tensor, size) -> torch.Tensor:
    if not tensor.is_contiguous():
        raise ValueError("reducescatter / allgather work in place and need a contiguous tensor")
    flat = tensor.view(-1)
    return flat if size is None else flat[: int(size)]


def shard_range(communicator, numel: int, element_size: int, rank=None) -> Tuple[int, int]:
    r = communicator.world_rank if rank is None else rank
    return shard_of(0, numel, r, communicator.world_size, max(1, 16 // element_size))


def reduce_scatter(communicator, tensor: torch.Tensor, size=None, op: str = "sum") -> Tuple[int, int]:
    flat = _flat(tensor, size)
    world = communicator.world_size
    mine = shard_range(communicator, flat.numel(), flat.element_size())
    if world == 1:
        return mine
    if flat.is_cuda:
        if not communicator.single_server:
            raise NotImplementedError("reducescatter across servers: use allreduce (hierarchical) and keep the shard")
        native = communicator._ensure_native()
        return native.reduce_scatter_(flat, op=op)
    rop = {"sum": dist.ReduceOp.SUM, "avg": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX}[op]
    for r in range(world):
        lo, hi = shard_range(communicator, flat.numel(), flat.element_size(), r)
        if hi > lo:
            dist.reduce(flat[lo:hi], dst=r, op=rop)
    if op == "avg":
        flat[mine[0]:mine[1]] /= world
    return mine


def all_gather(communicator, tensor: torch.Tensor, size=None) -> torch.Tensor:
    flat = _flat(tensor, size)
    world = communicator.world_size
    if world == 1:
        return tensor
    if flat.is_cuda:
        if not communicator.single_server:
            raise NotImplementedError("allgather across servers is not implemented")
        communicator._ensure_native().all_gather_(flat)
        return tensor
    for r in range(world):
        lo, hi = shard_range(communicator, flat.numel(), flat.element_size(), r)
        if hi > lo:
            dist.broadcast(flat[lo:hi], src=r)
    return tensor
