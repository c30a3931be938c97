"""Multi-round broadcast synthesis over an explicit link list (Blink / "fiddlelink"-style).

The reference carries a stand-alone research sketch for this: a Python-2 CVXPY boolean program that
schedules the broadcast of P partitions over an arc list in discrete rounds, minimising the number
of active time steps subject to forwarding ("a node can only send what it already holds"), one
partition per link per step, and no node receiving a partition twice
(/root/reference/gurobi/code-gen/cvxpy-broadcast-multi-round.py:44-153,174-233; inputs such as
``8-node-hgx.txt``). It is not wired into the library there. This is the same formulation on
``scipy.optimize.milp`` (HiGHS), wired in: ``schedule_broadcast`` returns the per-round transfers and
``to_strategy`` turns the first-arrival edges of every partition into a ``<trees>`` strategy (one tree
per partition) that the tree kernels execute.

    arcs = ring_arcs(4)                           # or parse_arc_file(".../8-node-hgx.txt")
    rounds = schedule_broadcast(4, arcs, root=0, partitions=2)
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from ..strategy.trees import Strategy, Tree

Arc = Tuple[int, int]


def ring_arcs(n: int, bidirectional: bool = True) -> List[Arc]:
    arcs = [(i, (i + 1) % n) for i in range(n)]
    if bidirectional:
        arcs += [((i + 1) % n, i) for i in range(n)]
    return arcs


def full_arcs(n: int) -> List[Arc]:
    return [(i, j) for i in range(n) for j in range(n) if i != j]


def parse_arc_file(path) -> Tuple[int, List[Arc]]:
    """Lines ``src dst`` (or ``src,dst``); returns (node count, arcs)."""
    arcs = []
    with open(path) as f:
        for ln in f:
            parts = ln.replace(",", " ").split()
            if len(parts) >= 2 and parts[0].lstrip("-").isdigit():
                arcs.append((int(parts[0]), int(parts[1])))
    n = 1 + max(max(a) for a in arcs)
    return n, arcs


def schedule_broadcast(n: int, arcs: Sequence[Arc], root: int = 0, partitions: int = 1,
                       max_rounds: Optional[int] = None, time_limit_s: float = 10.0):
    """Minimise the number of rounds in which any link is active.

    x[t, a, p] in {0,1}: partition p crosses arc a in round t.   has[t, v, p]: v holds p after round t.
    Returns a list over rounds of [(src, dst, partition), ...]."""
    from scipy.optimize import Bounds, LinearConstraint, milp
    from scipy.sparse import lil_matrix

    arcs = list(arcs)
    A, P = len(arcs), partitions
    T = max_rounds or (n - 1 + P - 1)
    nx = T * A * P
    nh = (T + 1) * n * P
    ny = T                                   # y[t] = round t is active
    nv = nx + nh + ny

    def X(t, a, p): return (t * A + a) * P + p
    def H(t, v, p): return nx + (t * n + v) * P + p
    def Y(t): return nx + nh + t

    rows = []
    for p in range(P):                       # initial holdings
        for v in range(n):
            rows.append(({H(0, v, p): 1.0}, 1.0 if v == root else 0.0, 1.0 if v == root else 0.0))
    for t in range(T):
        for a, (u, v) in enumerate(arcs):
            rows.append(({X(t, a, p): 1.0 for p in range(P)}, 0.0, 1.0))          # one partition per link per step
            for p in range(P):
                rows.append(({X(t, a, p): 1.0, H(t, u, p): -1.0}, -np.inf, 0.0))  # forward only what you hold
                rows.append(({X(t, a, p): 1.0, Y(t): -1.0}, -np.inf, 0.0))        # activity indicator
        for v in range(n):
            for p in range(P):
                inc = {X(t, a, p): 1.0 for a, (u, w) in enumerate(arcs) if w == v}
                c = dict(inc)
                c[H(t, v, p)] = 1.0
                c[H(t + 1, v, p)] = -1.0
                rows.append((c, 0.0, 0.0))                                        # has' = has + received
                # (has' <= 1 via bounds => never received twice / never received when already held)
    for v in range(n):
        for p in range(P):
            rows.append(({H(T, v, p): 1.0}, 1.0, 1.0))                            # everyone ends with everything
    M = lil_matrix((len(rows), nv))
    lo, hi = np.empty(len(rows)), np.empty(len(rows))
    for k, (c, l, h) in enumerate(rows):
        for j, w in c.items():
            M[k, j] = w
        lo[k], hi[k] = l, h
    cost = np.zeros(nv)
    for t in range(T):
        cost[Y(t)] = 1.0 + 1e-3 * t                                               # prefer early rounds
    cost[:nx] = 1e-5                                                              # and few transfers
    res = milp(c=cost, constraints=LinearConstraint(M.tocsr(), lo, hi), integrality=np.ones(nv),
               bounds=Bounds(np.zeros(nv), np.ones(nv)), options={"time_limit": time_limit_s, "disp": False})
    if res.x is None:
        raise RuntimeError(f"multi-round broadcast infeasible within {T} rounds: {res.message}")
    out = []
    for t in range(T):
        moves = [(arcs[a][0], arcs[a][1], p) for a in range(A) for p in range(P) if res.x[X(t, a, p)] > 0.5]
        if moves:
            out.append(moves)
    return out


def to_strategy(n: int, rounds, root: int, partitions: int, ips: Optional[Sequence[str]] = None) -> Strategy:
    """One tree per partition: parent(v) = the node v first received the partition from."""
    trees = []
    for p in range(partitions):
        t = Tree(root=root)
        for moves in rounds:
            for (u, v, q) in moves:
                if q == p and v != root and v not in t.parent:
                    t.parent[v] = u
                    t.children.setdefault(u, []).append(v)
        order: List[int] = []

        def dfs(x):
            order.append(x)
            for c in t.kids(x):
                dfs(c)
        dfs(root)
        t.nodes = order
        for x in order:
            t.ip[x] = ips[x] if ips else "127.0.0.1"
        trees.append(t)
    return Strategy(trees, {"policy": "multi-round-broadcast", "rounds": str(len(rounds))})
