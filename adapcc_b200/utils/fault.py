"""Fault / straggler injection for tests and experiments. The reference has none ("stragglers are
emulated only in analysis", ``--heter_alpha 2.7`` multiplies measured gaps,
/root/reference/units-test/get_wait_time.py:60,103); here delays and failures can be injected into a
live job so relay control and the heartbeat deadline are exercised for real.

    inj = FaultInjector.from_env(rank)         # ADAPCC_STRAGGLERS="6,7" ADAPCC_STRAGGLE_MS=250 ADAPCC_KILL="3@20"
    for step in ...:
        inj.before_backward(step)              # sleeps on straggler ranks, exits on the killed rank
"""
from __future__ import annotations

import os
import random
import sys
import time
from dataclasses import dataclass, field
from typing import Dict, Set


@dataclass
class FaultInjector:
    """Test / benchmark helper that makes chosen ranks slow (fixed delay, jitter, heterogeneity factor) or kills them
    at a given step, driven by ``ADAPCC_STRAGGLERS / ADAPCC_STRAGGLE_MS / ADAPCC_JITTER / ADAPCC_STRAGGLE_FROM /
    ADAPCC_KILL / ADAPCC_HETER_ALPHA`` — how the straggler and fault scenarios of the reference's evaluation are
    reproduced on one box."""

    rank: int
    stragglers: Set[int] = field(default_factory=set)
    straggle_ms: float = 0.0
    jitter: float = 0.0                      # +- fraction of straggle_ms, uniform
    from_step: int = 2
    kill_at: Dict[int, int] = field(default_factory=dict)     # rank -> step at which it dies
    heter_alpha: float = 1.0                 # multiplies the delay (the reference's analysis knob)
    log: list = field(default_factory=list)

    @classmethod
    def from_env(cls, rank: int) -> "FaultInjector":
        s = {int(x) for x in os.environ.get("ADAPCC_STRAGGLERS", "").split(",") if x.strip()}
        kill = {}
        for item in os.environ.get("ADAPCC_KILL", "").split(","):
            if "@" in item:
                r, st = item.split("@")
                kill[int(r)] = int(st)
        return cls(rank, s, float(os.environ.get("ADAPCC_STRAGGLE_MS", 0)), float(os.environ.get("ADAPCC_JITTER", 0)),
                   int(os.environ.get("ADAPCC_STRAGGLE_FROM", 2)), kill, float(os.environ.get("ADAPCC_HETER_ALPHA", 1)))

    def delay_s(self, step: int) -> float:
        if self.rank not in self.stragglers or step < self.from_step or self.straggle_ms <= 0:
            return 0.0
        d = self.straggle_ms * self.heter_alpha
        if self.jitter:
            d *= 1 + random.uniform(-self.jitter, self.jitter)
        return d / 1e3

    def before_backward(self, step: int) -> float:
        if self.kill_at.get(self.rank, -1) == step:
            print(f"[fault] rank {self.rank} exits at step {step}", flush=True)
            sys.stdout.flush()
            os._exit(17)
        d = self.delay_s(step)
        if d:
            time.sleep(d)
            self.log.append((step, d))
        return d
