"""MILP strategy solver (policy ``"gurobi"`` in the reference's vocabulary).

The reference builds a Gurobi model over root choice ``r_mg``, per-flow routing ``x_ijf``, chunk
size ``c_m``, aggregation control ``a_mj`` and link load ``N_mij`` — but it multiplies variables,
never calls ``model.optimize()`` and never writes XML (/root/reference/gurobi/solver.py:11-211;
SURVEY Appendix C.11). gurobipy is not installable here, so this is a linear reformulation that an
open solver (HiGHS through ``scipy.optimize.milp``) handles — an incumbent within ~1 s for 8 ranks (the
optimality proof of the highly symmetric switch case does not finish, so the solve is time-boxed) — and that
really emits the strategy:

  variables   x[m,i,j] in {0,1}   rank i's parent in tree m is j      (reduce edge i -> j)
              r[m,g]   in {0,1}   g is the root of tree m
              d[m,i]   in [0,n-1] depth of i in tree m                (MTZ, forbids cycles)
              D        >= d[m,i]  deepest node over all trees
              L        >= ingress time of any rank (all trees share a GPU's NVLink ingress port)
              E        >= egress time of any rank in the broadcast phase
  constraints sum_j x[m,i,j] = 1 - r[m,i];  sum_g r[m,g] = 1;  a rank roots at most ceil(M/n) trees;
              d[m,i] >= d[m,j] + 1 - n (1 - x[m,i,j]);  d[m,i] <= (n-1)(1 - r[m,i])
  objective   min  L + E + (alpha + c beta) * 2 D
              with L_j = sum_{m,i} x[m,i,j] * s_m * beta_ij   (s_m = slice bytes of tree m).

The chunk size is then chosen by a 1-D search of the analytic pipeline model over the solved trees
(the product ``num_chunks * T_bottleneck`` is what makes the reference's model non-linear).
Cross-server edges keep the ``ip`` attribute so an inter-node leg can be attached later.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np

from ..strategy.trees import Strategy, Tree
from .cost_model import LinkModel, best_chunk_bytes, strategy_time
from .partrees import DEFAULT_CHUNK


class SolverError(RuntimeError):
    pass


class Solver:
    """MILP synthesizer: parent choice, root choice, depth (MTZ) and ingress / egress load variables per tree, solved
    by HiGHS (``scipy.optimize.milp``) under a time limit, best of the incumbents by the cost model; emits the
    strategy XML and a chunk size. Same call signature as the reference's ``Solver.optimize``
    (/root/reference/gurobi/solver.py), whose Gurobi model is never solved."""

    def __init__(self, time_limit_s: float = 5.0, mip_rel_gap: float = 0.02, first_limit_s: float = 1.5):
        self.time_limit_s = time_limit_s
        self.first_limit_s = first_limit_s
        self.mip_rel_gap = mip_rel_gap
        self.last_status: Optional[str] = None
        self.last_objective: Optional[float] = None

    # -- model ---------------------------------------------------------------------------
    def solve(self, parallel_degree: int, transmission_bytes: float, bandwidth_graph, latency_graph,
              ip_table: Optional[Sequence[str]] = None, ranks: Optional[Sequence[int]] = None) -> Strategy:
        from scipy.optimize import Bounds, LinearConstraint, milp
        from scipy.sparse import lil_matrix

        lm = LinkModel(latency_graph, bandwidth_graph)
        R = list(range(lm.world)) if ranks is None else list(ranks)
        n = len(R)
        if n < 2:
            raise SolverError("need at least two ranks")
        M = max(1, min(parallel_degree, n))
        s_m = transmission_bytes / M
        edges = [(a, b) for a in range(n) for b in range(n) if a != b and lm.bw_gbs[R[a]][R[b]] > 0]
        if len(edges) < n - 1:
            raise SolverError("bandwidth graph is not connected")
        eidx = {e: k for k, e in enumerate(edges)}
        ne = len(edges)
        # variable layout
        off_x = 0
        off_r = off_x + M * ne
        off_d = off_r + M * n
        off_D = off_d + M * n
        off_L = off_D + 1
        off_E = off_L + 1
        nv = off_E + 1

        def X(m, a, b): return off_x + m * ne + eidx[(a, b)]
        def Rv(m, g): return off_r + m * n + g
        def Dv(m, i): return off_d + m * n + i

        rows: List[tuple] = []   # (coeff dict, lo, hi)

        def add(coeffs, lo, hi):
            rows.append((coeffs, lo, hi))

        for m in range(M):
            for i in range(n):       # exactly one parent unless root
                c = {X(m, i, j): 1.0 for j in range(n) if (i, j) in eidx}
                c[Rv(m, i)] = 1.0
                add(c, 1.0, 1.0)
                add({Dv(m, i): 1.0, Rv(m, i): float(n - 1)}, -np.inf, float(n - 1))  # root depth 0
                add({off_D: 1.0, Dv(m, i): -1.0}, 0.0, np.inf)
            add({Rv(m, g): 1.0 for g in range(n)}, 1.0, 1.0)
            for (i, j) in edges:     # MTZ: d_i >= d_j + 1 - n(1 - x_ij)
                add({Dv(m, i): 1.0, Dv(m, j): -1.0, X(m, i, j): -float(n)}, 1.0 - n, np.inf)
        cap = math.ceil(M / n)
        for g in range(n):           # spread the roots
            add({Rv(m, g): 1.0 for m in range(M)}, 0.0, float(cap))
        if lm.is_uniform() and len({lm.bw_gbs[R[a]][R[b]] > 0 for a in range(n) for b in range(n) if a != b}) == 1:
            # symmetry breaking on a switch (all links alike): which rank roots which tree is arbitrary, so pin
            # tree m's root to rank m mod n instead of letting branch-and-bound explore the n!/(n-M)! relabelings
            for m in range(M):
                add({Rv(m, m % n): 1.0}, 1.0, 1.0)
        scale = 1e6                  # work in microseconds for conditioning
        for j in range(n):
            cin = {X(m, i, j): -s_m * lm.beta(R[i], R[j]) * scale for m in range(M) for i in range(n)
                   if (i, j) in eidx}
            cin[off_L] = 1.0
            add(cin, 0.0, np.inf)    # L >= ingress_j
            cout = {X(m, i, j): -s_m * lm.beta(R[j], R[i]) * scale for m in range(M) for i in range(n)
                    if (i, j) in eidx and lm.bw_gbs[R[j]][R[i]] > 0}
            cout[off_E] = 1.0
            add(cout, 0.0, np.inf)   # E >= egress_j (broadcast pushes to children)

        A = lil_matrix((len(rows), nv))
        lo = np.empty(len(rows))
        hi = np.empty(len(rows))
        for k, (c, l, h) in enumerate(rows):
            for v, w in c.items():
                A[k, v] = w
            lo[k], hi[k] = l, h

        chunk0 = min(float(DEFAULT_CHUNK), max(16.0, s_m))
        hop = (lm.mean_alpha(R) + chunk0 / (lm.min_bw(R) * 1e9)) * scale
        cost = np.zeros(nv)
        cost[off_L] = 1.0
        cost[off_E] = 1.0
        cost[off_D] = 2.0 * hop
        # tiny tie-break towards low-latency edges
        for m in range(M):
            for (i, j) in edges:
                cost[X(m, i, j)] = 1e-3 * lm.alpha(R[i], R[j]) * scale

        integrality = np.zeros(nv)
        integrality[off_x:off_d] = 1
        lb = np.zeros(nv)
        ub = np.full(nv, np.inf)
        ub[off_x:off_d] = 1.0
        ub[off_d:off_D] = float(n - 1)
        # The incumbent usually appears within a second and then only the (symmetric) optimality proof runs until the
        # limit, so the first attempt is short; the long one only happens when nothing feasible was found yet.
        res = None
        for limit in (min(self.time_limit_s, self.first_limit_s), self.time_limit_s):
            res = milp(c=cost, constraints=LinearConstraint(A.tocsr(), lo, hi), integrality=integrality,
                       bounds=Bounds(lb, ub),
                       options={"time_limit": limit, "mip_rel_gap": self.mip_rel_gap, "disp": False})
            if res.x is not None or limit >= self.time_limit_s:
                break
        self.last_status = str(res.message)
        if res.x is None:
            raise SolverError(f"MILP found no solution: {res.message}")
        self.last_objective = float(res.fun)
        x = res.x
        trees: List[Tree] = []
        for m in range(M):
            root = max(range(n), key=lambda g: x[Rv(m, g)])
            t = Tree(root=R[root])
            par = {}
            for (i, j) in edges:
                if x[X(m, i, j)] > 0.5:
                    par[i] = j
            kids = {}
            for i, j in par.items():
                kids.setdefault(j, []).append(i)
            order: List[int] = []

            def dfs(a: int) -> None:
                order.append(a)
                for c in sorted(kids.get(a, [])):
                    t.parent[R[c]] = R[a]
                    t.children.setdefault(R[a], []).append(R[c])
                    dfs(c)
            dfs(root)
            if len(order) != n:
                raise SolverError("solution is not a spanning tree (solver tolerance)")
            t.nodes = [R[a] for a in order]
            for a in order:
                t.ip[R[a]] = ip_table[R[a]] if ip_table else "127.0.0.1"
            trees.append(t)
        return Strategy(trees, {"policy": "milp"})

    # -- reference signature -------------------------------------------------------------
    def optimize(self, prim, parallel_degree, transmission_size, bandwidth_graph, latency_graph,
                 strategy_file, ip_table: Optional[Sequence[str]] = None):
        """Same arguments as /root/reference/gurobi/solver.py:11-12 (``transmission_size`` in fp32
        elements). Solves, writes ``strategy_file`` and returns the chunk size in bytes."""
        total = float(transmission_size) * 4.0
        s = self.solve(parallel_degree, total, bandwidth_graph, latency_graph, ip_table)
        lm = LinkModel(latency_graph, bandwidth_graph)
        chunk = best_chunk_bytes(s, lm, total)
        s.attrs.update({"prim": str(prim), "chunk": str(chunk),
                        "est_us": f"{strategy_time(s, lm, total, chunk) * 1e6:.1f}"})
        if strategy_file:
            s.save(strategy_file, compact=True)
        return chunk
