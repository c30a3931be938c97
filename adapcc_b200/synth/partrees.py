"""ParTrees: the reference's default synthesis policy, rebuilt.

Reference behaviour (/root/reference/gurobi/trees.py:110-152): one node per server (keyed by its
local rank 0) carrying the measured bandwidth/latency to the next server, sorted by
bandwidth x latency product (descending); ``parallel_degree`` binary trees across servers, each a
rotation of that list; a CHAIN inside every server; emitted as ``<trees><root><gpu>...`` XML;
returns the default chunk size (4 MiB).

Additions for a single NVSwitch box (where the reference degenerates to ONE 8-GPU chain because
``parallel_degree = min(#servers, degree)``): ``intra_policy`` can be ``"chain"`` (reference),
``"binary"`` or ``"star"``, and on a single server the roots are rotated across the GPUs so
``parallel_degree`` trees exist and every NVLink port carries traffic.
"""
from __future__ import annotations

import copy
from typing import Dict, List, Optional, Sequence

from ..strategy.trees import Strategy, Tree
from .cost_model import LinkModel, best_chunk_bytes

DEFAULT_CHUNK = 4 * 1024 * 1024


class _Node:
    def __init__(self, world_rank: int, ip: str, bandwidth: float, latency: float):
        self.world_rank, self.ip = world_rank, ip
        self.bandwidth, self.latency = bandwidth, latency
        self.bdp = bandwidth * latency
        self.left: Optional["_Node"] = None
        self.right: Optional["_Node"] = None


def _heap_link(nodes: List[_Node]) -> _Node:
    for i, nd in enumerate(nodes):
        l, r = 2 * i + 1, 2 * i + 2
        nd.left = nodes[l] if l < len(nodes) else None
        nd.right = nodes[r] if r < len(nodes) else None
    return nodes[0]


def _intra(tree: Tree, head: int, members: Sequence[int], ip: str, policy: str) -> None:
    """Attach the server's other GPUs under its head GPU."""
    rest = [g for g in members if g != head]
    if policy == "chain":
        prev = head
        for g in rest:
            _attach(tree, g, prev, ip)
            prev = g
    elif policy == "star":
        for g in rest:
            _attach(tree, g, head, ip)
    elif policy == "binary":
        order = [head] + rest
        for i, g in enumerate(order[1:], start=1):
            _attach(tree, g, order[(i - 1) // 2], ip)
    else:
        raise ValueError(f"unknown intra policy {policy!r}")


def _attach(tree: Tree, child: int, parent: int, ip: str) -> None:
    tree.parent[child] = parent
    tree.children.setdefault(parent, []).append(child)
    tree.ip[child] = ip


def _finalise(tree: Tree) -> Tree:
    order: List[int] = []

    def dfs(x: int) -> None:
        order.append(x)
        for c in tree.kids(x):
            dfs(c)
    dfs(tree.root)
    tree.nodes = order
    return tree


class ParTrees:
    """The reference's parallel-trees heuristic (/root/reference/gurobi/trees.py:5-152): servers chained / paired by
    the parity rule, GPUs inside a server arranged by ``intra_policy`` (chain | binary | star), one tree per
    parallel transmission with rotated roots on a single NVSwitch server."""

    def __init__(self, intra_policy: str = "chain", rotate_single_server: bool = True):
        self.intra_policy = intra_policy
        self.rotate_single_server = rotate_single_server

    def build(self, ip_table: Sequence[str], local_rank0_list: Sequence[int], parallel_degree: int,
              bandwidth_graph, latency_graph) -> Strategy:
        world = len(ip_table)
        groups: Dict[int, List[int]] = {}
        nodes: List[_Node] = []
        for r0 in local_rank0_list:
            g, idx = [], r0
            while idx < world and ip_table[idx] == ip_table[r0]:
                g.append(idx)
                idx += 1
            groups[r0] = g
            nxt = (r0 + len(g)) % world
            nodes.append(_Node(r0, ip_table[r0], bandwidth_graph[r0][nxt], latency_graph[r0][nxt]))
        nodes.sort(key=lambda nd: nd.bdp, reverse=True)

        trees: List[Tree] = []
        if len(nodes) == 1 and self.rotate_single_server:
            members = groups[nodes[0].world_rank]
            degree = max(1, min(parallel_degree, len(members)))
            step = max(1, len(members) // degree)
            for k in range(degree):
                rot = members[k * step:] + members[:k * step]
                t = Tree(root=rot[0])
                t.ip[rot[0]] = nodes[0].ip
                _intra(t, rot[0], rot, nodes[0].ip, self.intra_policy)
                trees.append(_finalise(t))
            return Strategy(trees, {"policy": "par-trees", "intra": self.intra_policy})

        degree = max(1, min(len(local_rank0_list), parallel_degree))
        order = list(nodes)
        for k in range(degree):
            if k > 0:
                order = order[1:] + order[:1]
            top = _heap_link(copy.deepcopy(order))
            t = Tree(root=top.world_rank)
            t.ip[top.world_rank] = top.ip

            def rec(nd: _Node) -> None:
                _intra(t, nd.world_rank, groups[nd.world_rank], nd.ip, self.intra_policy)
                for ch in (nd.left, nd.right):
                    if ch is not None:
                        _attach(t, ch.world_rank, nd.world_rank, ch.ip)
                        rec(ch)
            rec(top)
            trees.append(_finalise(t))
        return Strategy(trees, {"policy": "par-trees", "intra": self.intra_policy})

    def optimize(self, ip_table, local_rank0_list, prim, parallel_degree, transmission_size,
                 bandwidth_graph, latency_graph, strategy_file) -> int:
        """Reference signature (/root/reference/gurobi/trees.py:114-116): writes the XML, returns the
        chunk size in bytes."""
        s = self.build(ip_table, local_rank0_list, parallel_degree, bandwidth_graph, latency_graph)
        chunk = DEFAULT_CHUNK
        try:
            lm = LinkModel(latency_graph, bandwidth_graph)
            if any(b > 0 for row in bandwidth_graph for b in row):
                chunk = best_chunk_bytes(s, lm, float(transmission_size) * 4)
        except Exception:
            chunk = DEFAULT_CHUNK
        s.attrs["chunk"] = str(chunk)
        s.attrs["prim"] = str(prim)
        if strategy_file:
            s.save(strategy_file, compact=True)
        return chunk
