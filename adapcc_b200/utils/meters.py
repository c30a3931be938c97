"""Observability helpers. The reference only prints (``[Rank n]...`` lines, AverageMeter /
ProgressMeter in its ImageNet scripts, /root/reference/models/image-classification/main_elastic.py:515-554);
here the same meters plus a structured JSON-lines metrics sink and a per-op bandwidth report from
the library itself (which the reference never produced)."""
from __future__ import annotations

import json
import os
import time
from typing import List, Optional


class AverageMeter:
    """Running value / average of a metric, printed as ``name val (avg)`` (the meters of the reference's ImageNet
    scripts, /root/reference/models/image-classification/main_elastic.py:515-554)."""

    def __init__(self, name: str, fmt: str = ":f"):
        self.name, self.fmt = name, fmt
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = 0.0
        self.count = 0

    def update(self, val: float, n: int = 1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / max(1, self.count)

    def __str__(self):
        return ("{name} {val" + self.fmt + "} ({avg" + self.fmt + "})").format(**self.__dict__)


class ProgressMeter:
    """One progress line per call: ``prefix[batch/total]`` followed by the meters, tab separated (same layout as the
    reference's trainers, so its log processors keep working)."""

    def __init__(self, num_batches: int, meters: List[AverageMeter], prefix: str = ""):
        n = len(str(num_batches))
        self.fmt = "[{:" + str(n) + "d}/" + ("{:" + str(n) + "d}").format(num_batches) + "]"
        self.meters, self.prefix = meters, prefix

    def display(self, batch: int) -> str:
        line = "\t".join([self.prefix + self.fmt.format(batch)] + [str(m) for m in self.meters])
        print(line, flush=True)
        return line


class MetricsSink:
    """Append-only JSON-lines file (one record per event), rank-tagged."""

    def __init__(self, path: Optional[str], rank: int = 0):
        self.path, self.rank = path, rank
        if path:
            os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)

    def emit(self, event: str, **fields):
        if not self.path:
            return
        rec = {"ts": time.time(), "rank": self.rank, "event": event, **fields}
        with open(self.path, "a") as f:
            f.write(json.dumps(rec) + "\n")


def busbw_gbs(nbytes: int, seconds: float, world: int, prim: str = "allreduce") -> float:
    """Bus bandwidth with the nccl-tests correction factors (/root/reference/nccl-perf/benchmark/PERFORMANCE.md:33-142):
    all-reduce 2(n-1)/n, all-gather / reduce-scatter / all-to-all (n-1)/n of the TOTAL buffer size, reduce and
    broadcast 1."""
    n = max(1, world)
    factor = {"allreduce": 2 * (n - 1) / n, "allgather": (n - 1) / n, "reducescatter": (n - 1) / n,
              "alltoall": (n - 1) / n}.get(prim, 1.0)
    return nbytes / max(seconds, 1e-12) * factor / 1e9
