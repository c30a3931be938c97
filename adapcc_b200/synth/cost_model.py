"""alpha-beta cost model over the profiler's measured latency / bandwidth matrices.

Units follow the profile dump (/root/reference/csrc/profile.cu:336-357): latency in microseconds,
bandwidth in GB/s. The model feeds (a) the tree synthesizers and (b) the per-message algorithm
choice (one-shot / two-shot / NVLS / tree) — "variant picked per chunk from the profiler's measured
link bandwidth" (BASELINE.json north star).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

from ..strategy.trees import Strategy, Tree

GB = 1e9


@dataclass
class LinkModel:
    alpha_us: List[List[float]]     # [src][dst] one-way latency, us
    bw_gbs: List[List[float]]       # [src][dst] bandwidth, GB/s

    @classmethod
    def uniform(cls, world: int, alpha_us: float = 2.0, bw_gbs: float = 700.0) -> "LinkModel":
        return cls([[0.0 if i == j else alpha_us for j in range(world)] for i in range(world)],
                   [[0.0 if i == j else bw_gbs for j in range(world)] for i in range(world)])

    @property
    def world(self) -> int:
        return len(self.bw_gbs)

    def beta(self, i: int, j: int) -> float:
        """seconds per byte on i -> j (inf when unmeasured)."""
        b = self.bw_gbs[i][j]
        return 1.0 / (b * GB) if b > 0 else float("inf")

    def alpha(self, i: int, j: int) -> float:
        return self.alpha_us[i][j] * 1e-6

    def mean_alpha(self, ranks: Optional[Sequence[int]] = None) -> float:
        rs = list(range(self.world)) if ranks is None else list(ranks)
        v = [self.alpha(i, j) for i in rs for j in rs if i != j and self.bw_gbs[i][j] > 0]
        return sum(v) / len(v) if v else 2e-6

    def min_bw(self, ranks: Optional[Sequence[int]] = None) -> float:
        rs = list(range(self.world)) if ranks is None else list(ranks)
        v = [self.bw_gbs[i][j] for i in rs for j in rs if i != j and self.bw_gbs[i][j] > 0]
        return min(v) if v else 1.0

    def is_uniform(self, tol: float = 0.25) -> bool:
        v = [self.bw_gbs[i][j] for i in range(self.world) for j in range(self.world)
             if i != j and self.bw_gbs[i][j] > 0]
        return bool(v) and (max(v) - min(v)) <= tol * max(v)


def tree_time(tree: Tree, lm: LinkModel, slice_bytes: float, chunk_bytes: float, bcast: bool = True,
              ingress_share: Optional[Dict[int, float]] = None) -> float:
    """Pipelined completion time of one reduction(+broadcast) tree over its slice.

    Per chunk a parent pulls from all its children through ONE ingress port, so its step time is
    alpha + sum_children chunk*beta; the pipeline runs at the slowest node's step time and fills in
    depth steps (reduce up, broadcast down)."""
    if not tree.nodes or slice_bytes <= 0:
        return 0.0
    chunk = max(16.0, min(chunk_bytes, slice_bytes))
    n_chunks = max(1.0, slice_bytes / chunk)
    step_max = 0.0

    def step(x: int) -> float:
        kids = tree.kids(x)
        if not kids:
            return 0.0
        share = (ingress_share or {}).get(x, 1.0)
        return max(lm.alpha(c, x) for c in kids) + sum(chunk * lm.beta(c, x) for c in kids) * share

    def fill(x: int) -> float:
        kids = tree.kids(x)
        if not kids:
            return 0.0
        return step(x) + max(fill(c) for c in kids)

    for x in tree.nodes:
        step_max = max(step_max, step(x))
    t = fill(tree.root) + (n_chunks - 1) * step_max
    if bcast:
        # same edges backwards; a child's pull is one flow on its own ingress
        def bfill(x: int) -> float:
            kids = tree.kids(x)
            if not kids:
                return 0.0
            return max(lm.alpha(x, c) + chunk * lm.beta(x, c) + bfill(c) for c in kids)
        egress = max((len(tree.kids(x)) for x in tree.nodes), default=1)
        bstep = max((lm.alpha(x, c) + chunk * lm.beta(x, c) * max(1, len(tree.kids(x)))
                     for x in tree.nodes for c in tree.kids(x)), default=0.0)
        t = max(t, fill(tree.root)) + bfill(tree.root) + max(0.0, (n_chunks - 1) * (bstep - step_max))
        _ = egress
    return t


def strategy_time(strategy: Strategy, lm: LinkModel, total_bytes: float, chunk_bytes: float,
                  bcast: bool = True) -> float:
    """Trees run concurrently on disjoint slices; nodes that are parents in several trees share
    their ingress port between them."""
    nt = len(strategy.trees)
    if nt == 0:
        return float("inf")
    load: Dict[int, int] = {}
    for t in strategy.trees:
        for x in t.nodes:
            if t.kids(x):
                load[x] = load.get(x, 0) + 1
    per = total_bytes / nt
    return max(tree_time(t, lm, per, chunk_bytes, bcast, {x: float(k) for x, k in load.items()})
               for t in strategy.trees)


def direct_times(lm: LinkModel, nbytes: float, ranks: Optional[Sequence[int]] = None,
                 nvls: bool = True, nvls_bw_gbs: Optional[float] = None) -> Dict[str, float]:
    """Closed-form estimates for the switch-topology algorithms (seconds)."""
    rs = list(range(lm.world)) if ranks is None else list(ranks)
    n = max(2, len(rs))
    a = lm.mean_alpha(rs)
    beta = 1.0 / (lm.min_bw(rs) * GB)
    out = {
        # one barrier + every rank pulls (n-1) windows + exit barrier
        "one_shot": 3 * a + (n - 1) * nbytes * beta,
        # barrier, pull slice from n-1 peers, push to n-1 peers, barrier
        "two_shot": 4 * a + 2 * (n - 1) / n * nbytes * beta,
    }
    if nvls:
        mb = 1.0 / ((nvls_bw_gbs or lm.min_bw(rs)) * GB)
        # the switch reduces on the way in (S/n per rank) and replicates on the way out (S)
        out["nvls"] = 4 * a + (1.0 + 1.0 / n) * nbytes * mb
    return out


def pick_algorithm(lm: LinkModel, nbytes: float, ranks=None, nvls: bool = True,
                   strategy: Optional[Strategy] = None, chunk_bytes: float = 1 << 20) -> str:
    t = direct_times(lm, nbytes, ranks, nvls)
    if strategy is not None:
        t["tree"] = strategy_time(strategy, lm, nbytes, chunk_bytes)
    return min(t, key=t.get)


def crossover_bytes(lm: LinkModel, a: str, b: str, ranks=None, nvls: bool = True) -> int:
    """Smallest power-of-two message size where algorithm ``b`` beats ``a``."""
    for p in range(8, 34):
        t = direct_times(lm, float(1 << p), ranks, nvls)
        if a in t and b in t and t[b] < t[a]:
            return 1 << p
    return 1 << 34


def best_chunk_bytes(strategy: Strategy, lm: LinkModel, total_bytes: float,
                     candidates=(16 << 10, 64 << 10, 256 << 10, 1 << 20, 4 << 20)) -> int:
    return min(candidates, key=lambda c: strategy_time(strategy, lm, total_bytes, c))
