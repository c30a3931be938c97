"""Checkpoint / resume helpers for the elastic example. Not part of AdapCC proper — the reference has
them only in its torchelastic ImageNet script (``State.capture_snapshot/apply_snapshot``, atomic save
= tmp + rename, "broadcast the newest checkpoint from the max-epoch rank over a temporary gloo
group", /root/reference/models/image-classification/main_elastic.py:188-237,306-409). The
communication layer's own durable state is files (topology/*.xml, topo_profile_*, strategy/*.xml);
restarting with ``entry_point=-1`` reuses the last strategy."""
from __future__ import annotations

import io
import os
import tempfile
from typing import Any, Dict, Optional

import torch
import torch.distributed as dist


class State:
    """Everything a worker needs to resume (model, optimizer, epoch, step, extras) with atomic save / load and in-
    memory snapshots that can be broadcast to restarted workers — the ``State`` object of the reference's elastic
    ImageNet example (/root/reference/models/image-classification/main_elastic.py:188-305)."""

    def __init__(self, model, optimizer, epoch: int = -1, step: int = 0, extra: Optional[Dict[str, Any]] = None):
        self.model, self.optimizer, self.epoch, self.step = model, optimizer, epoch, step
        self.extra = extra or {}

    def capture_snapshot(self) -> Dict[str, Any]:
        return {"epoch": self.epoch, "step": self.step, "model": self.model.state_dict(),
                "optimizer": self.optimizer.state_dict() if self.optimizer is not None else None, "extra": self.extra}

    def apply_snapshot(self, snap: Dict[str, Any], device=None) -> None:
        self.epoch, self.step, self.extra = snap["epoch"], snap.get("step", 0), snap.get("extra", {})
        self.model.load_state_dict(snap["model"])
        if self.optimizer is not None and snap.get("optimizer") is not None:
            self.optimizer.load_state_dict(snap["optimizer"])

    def save(self, path: str) -> None:
        """Atomic: write to a temp file in the same directory, fsync, rename."""
        d = os.path.dirname(os.path.abspath(path))
        os.makedirs(d, exist_ok=True)
        fd, tmp = tempfile.mkstemp(dir=d, suffix=".tmp")
        with os.fdopen(fd, "wb") as f:
            torch.save(self.capture_snapshot(), f)
            f.flush()
            os.fsync(f.fileno())
        os.replace(tmp, path)

    def load(self, path: str, device="cpu") -> bool:
        if not os.path.isfile(path):
            return False
        self.apply_snapshot(torch.load(path, map_location=device, weights_only=False), device)
        return True


def broadcast_newest(state: State, group=None) -> int:
    """Every rank may hold a different (or no) checkpoint after an elastic restart: find the rank
    with the highest (epoch, step) and broadcast its snapshot to everyone. Returns the source rank."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = torch.tensor([state.epoch, state.step], dtype=torch.long)
    allv = [torch.zeros(2, dtype=torch.long) for _ in range(world)]
    dist.all_gather(allv, mine, group=group)
    src = max(range(world), key=lambda r: (int(allv[r][0]), int(allv[r][1]), -r))
    if int(allv[src][0]) < 0 and int(allv[src][1]) == 0:
        return src                                    # nobody has a checkpoint
    buf = io.BytesIO()
    if rank == src:
        torch.save(state.capture_snapshot(), buf)
    obj = [buf.getvalue() if rank == src else None]
    dist.broadcast_object_list(obj, src=src, group=group)
    if rank != src:
        state.apply_snapshot(torch.load(io.BytesIO(obj[0]), map_location="cpu", weights_only=False))
    return src
