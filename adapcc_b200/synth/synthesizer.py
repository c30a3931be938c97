"""Strategy synthesizer facade — API parity with /root/reference/gurobi/synthesizer.py:5-62.

``policy``: ``"par-trees"`` (reference default), ``"gurobi"``/``"milp"`` (the MILP, solved with
HiGHS), or ``"auto"`` (evaluate both plus the NVSwitch shapes with the cost model and keep the
fastest). ``generate_strategy`` writes the strategy XML and returns the chunk size in bytes.
"""
from __future__ import annotations

from typing import List, Optional

from ..strategy.trees import Strategy, make_strategy
from .cost_model import LinkModel, best_chunk_bytes, strategy_time
from .partrees import ParTrees
from .solver import Solver, SolverError


class Synthesizer:
    """Front end of strategy synthesis with the reference's constructor, setters and ``generate_strategy(prim) -> chunk
    bytes`` (/root/reference/gurobi/synthesizer.py:5-62); policies: ``par-trees`` (heuristic), ``milp`` / ``gurobi``
    (HiGHS), ``auto`` (cheaper of the two under the cost model)."""

    def __init__(self, strategy_file, ip_table=None, parallel_degree=4, size=10 * (10 ** 6),
                 bandwidth_graph=None, latency_graph=None, policy="par-trees", intra_policy="chain"):
        self.strategy_file = strategy_file
        self.ip_table = list(ip_table) if ip_table is not None else []
        self.local_rank0_list = self._get_local_rank0_list()
        self.parallel_degree = parallel_degree
        self.size = size
        self.bandwidth_graph = bandwidth_graph
        self.latency_graph = latency_graph
        self.policy = policy
        self.intra_policy = intra_policy
        self.last_strategy: Optional[Strategy] = None
        self.last_report = {}

    # ---- setters (reference names) -------------------------------------------------------
    def set_parallel_degree(self, parallel_degree):
        self.parallel_degree = parallel_degree

    def set_transmission_size(self, size):
        self.size = size

    def set_bandwidth_graph(self, graph):
        self.bandwidth_graph = graph

    def set_latency_graph(self, graph):
        self.latency_graph = graph

    def set_ip_info(self, ip_table):
        self.ip_table = list(ip_table)
        self.local_rank0_list = self._get_local_rank0_list()

    def _get_local_rank0_list(self) -> List[int]:
        out, seen = [], set()
        for r, ip in enumerate(self.ip_table):
            if ip not in seen:
                out.append(r)
                seen.add(ip)
        return out

    # ---- synthesis -----------------------------------------------------------------------
    def _graphs(self):
        w = len(self.ip_table)
        bw = self.bandwidth_graph or [[0.0 if i == j else 1.0 for j in range(w)] for i in range(w)]
        lat = self.latency_graph or [[0.0 if i == j else 1.0 for j in range(w)] for i in range(w)]
        return bw, lat

    def generate_strategy(self, prim):
        if prim not in ("reduce", "broadcast", "alltoall"):
            print("prim not within the formulation scope.")
            return None
        bw, lat = self._graphs()
        if self.policy in ("gurobi", "milp"):
            try:
                chunk = Solver().optimize(prim, self.parallel_degree, self.size, bw, lat, self.strategy_file,
                                          ip_table=self.ip_table)
                self.last_strategy = Strategy.from_file(self.strategy_file)
                return chunk
            except (SolverError, ImportError) as e:   # fall back like a production system must
                print(f"[synthesizer] MILP unavailable ({e}); using par-trees")
        if self.policy == "auto":
            return self._auto(prim, bw, lat)
        pt = ParTrees(intra_policy=self.intra_policy)
        chunk = pt.optimize(self.ip_table, self.local_rank0_list, prim, self.parallel_degree, self.size, bw, lat,
                            self.strategy_file)
        self.last_strategy = Strategy.from_file(self.strategy_file)
        return chunk

    def _auto(self, prim, bw, lat):
        lm = LinkModel(lat, bw)
        total = float(self.size) * 4.0
        world = len(self.ip_table)
        cands = {}
        for intra in ("chain", "binary", "star"):
            cands[f"par-trees/{intra}"] = ParTrees(intra).build(self.ip_table, self.local_rank0_list,
                                                               self.parallel_degree, bw, lat)
        if len(self.local_rank0_list) == 1:
            for shape in ("binary", "star"):
                cands[f"rotated/{shape}"] = make_strategy(world, self.parallel_degree, shape, self.ip_table)
        try:
            cands["milp"] = Solver(time_limit_s=3.0).solve(self.parallel_degree, total, bw, lat, self.ip_table)
        except Exception as e:  # noqa: BLE001
            self.last_report["milp_error"] = str(e)
        scored = {}
        for name, s in cands.items():
            c = best_chunk_bytes(s, lm, total)
            scored[name] = (strategy_time(s, lm, total, c), c)
        best = min(scored, key=lambda k: scored[k][0])
        self.last_report.update({k: {"est_us": v[0] * 1e6, "chunk": v[1]} for k, v in scored.items()})
        self.last_report["chosen"] = best
        s = cands[best]
        s.attrs.update({"policy": best, "chunk": str(scored[best][1]), "prim": str(prim)})
        s.save(self.strategy_file, compact=True)
        self.last_strategy = s
        return scored[best][1]
