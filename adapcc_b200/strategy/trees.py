"""Strategy model: a set of parallel reduction/broadcast trees ("transmissions").

Schema (/root/reference/strategy/*.xml): ``<trees>`` holds one ``<root id ip>``
per tree; nested ``<gpu id ip>`` elements are children (a child sends to its parent in the reduce
phase, the broadcast runs the same edges backwards). Tree *t* (document order) owns slice *t* of
the tensor. Cross-node edge <=> the ``ip`` strings differ.

Beyond the reference: validation, contraction of absent ranks (``world`` smaller than the file's
rank set, needed for BASELINE config 1: strategy/4.xml at world_size=2) and of inactive ranks
(NVSwitch "bypass" relay mode), and built-in generators for ring/direct/binary shapes.
"""
from __future__ import annotations

import os

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

from . import xmlio


class StrategyError(ValueError):
    pass


@dataclass
class Tree:
    """One reduction / broadcast tree of a strategy: ``parent`` map over world ranks plus each node's ``ip`` (a cross-
    server edge = different ips); child → parent edges reduce, the reverse edges broadcast
    (/root/reference/csrc/allreduce.cu:52-104)."""

    root: int = -1
    nodes: List[int] = field(default_factory=list)            # DFS pre-order (document order)
    parent: Dict[int, int] = field(default_factory=dict)      # child -> parent
    children: Dict[int, List[int]] = field(default_factory=dict)
    ip: Dict[int, str] = field(default_factory=dict)

    def kids(self, r: int) -> List[int]:
        return self.children.get(r, [])

    def depth(self) -> int:
        def rec(x: int) -> int:
            return 1 + max((rec(c) for c in self.kids(x)), default=0)
        return rec(self.root) if self.root >= 0 else 0

    def contract(self, keep: Callable[[int], bool]) -> "Tree":
        """Drop ranks failing ``keep``; orphans re-attach to their nearest kept ancestor. If the
        root is dropped the first kept top-level node becomes the root and adopts the others."""
        o = Tree()
        tops: List[int] = []
        for x in self.nodes:
            if not keep(x):
                continue
            o.nodes.append(x)
            o.ip[x] = self.ip.get(x, "")
            a, found = x, -1
            while a in self.parent:
                a = self.parent[a]
                if keep(a):
                    found = a
                    break
            if found >= 0:
                o.parent[x] = found
                o.children.setdefault(found, []).append(x)
            else:
                tops.append(x)
        if not tops:
            return o
        o.root = self.root if keep(self.root) else tops[0]
        for x in tops:
            if x != o.root:
                o.parent[x] = o.root
                o.children.setdefault(o.root, []).append(x)
        return o

    def to_node(self) -> xmlio.Node:
        def rec(x: int, tag: str) -> xmlio.Node:
            n = xmlio.Node(tag, {"id": str(x), "ip": self.ip.get(x, "")})
            n.children = [rec(c, "gpu") for c in self.kids(x)]
            return n
        return rec(self.root, "root")

    def edges(self) -> List[tuple]:
        return [(c, p) for c, p in self.parent.items()]


@dataclass
class Strategy:
    """A strategy file in memory: the parallel trees (tree *t* owns slice *t* of every tensor), the ``<trees>``
    attributes (chunk size, forced algorithm, the per-message algorithm plan) and validation / pruning to a world
    size — the XML schema of /root/reference/strategy/*.xml."""

    trees: List[Tree] = field(default_factory=list)
    attrs: Dict[str, str] = field(default_factory=dict)       # optional <trees algo=".." chunk="..">

    # ---- parsing -------------------------------------------------------------------------
    @classmethod
    def from_xml(cls, text: str, world: Optional[int] = None, max_trees: int = 8) -> "Strategy":
        doc = xmlio.parse(text)
        if doc.tag != "trees":
            raise StrategyError(f"root element is <{doc.tag}>, expected <trees>")
        s = cls(attrs=dict(doc.attrs))
        for r in doc.find_all("root"):
            t = Tree()
            _add_subtree(r, -1, t)
            if world is not None and world > 0:
                t = t.contract(lambda x: x < world)
            if t.root >= 0:
                s.trees.append(t)
        if not s.trees:
            raise StrategyError("no usable <root> tree")
        if len(s.trees) > max_trees:
            raise StrategyError(f"more than {max_trees} trees")
        return s

    @classmethod
    def from_file(cls, path, world: Optional[int] = None) -> "Strategy":
        with open(path, "r") as f:
            return cls.from_xml(f.read(), world)

    # ---- writing -------------------------------------------------------------------------
    def to_xml(self, compact: bool = False) -> str:
        """``compact=True``: one transmission tree per line with a describing comment (what the shipped
        sample strategies use); default: indented, one element per line."""
        if not compact:
            doc = xmlio.Node("trees", dict(self.attrs))
            doc.children = [t.to_node() for t in self.trees]
            return xmlio.dumps(doc)

        def rec(t: Tree, x: int, tag: str) -> str:
            a = f'ip="{t.ip.get(x, "")}" id="{x}"'
            kids = t.kids(x)
            if not kids:
                return f"<{tag} {a}/>"
            return f"<{tag} {a}>" + "".join(rec(t, c, "gpu") for c in kids) + f"</{tag}>"

        attrs = "".join(f' {k}="{v}"' for k, v in self.attrs.items())
        out = ['<?xml version="1.0" encoding="utf-8"?>', f"<trees{attrs}>"]
        for i, t in enumerate(self.trees):
            out.append(f"  <!-- tree {i}: slice {i} of the tensor, root rank {t.root}, depth {t.depth()}, "
                       f"{len(t.nodes)} ranks -->")
            out.append("  " + rec(t, t.root, "root"))
        out.append("</trees>")
        return "\n".join(out) + "\n"

    def save(self, path, compact: bool = False) -> None:
        parent = os.path.dirname(os.path.abspath(path))
        os.makedirs(parent, exist_ok=True)            # e.g. ./strategy/ when training starts outside the repo root
        with open(path, "w") as f:
            f.write(self.to_xml(compact))

    # ---- queries -------------------------------------------------------------------------
    def ranks(self) -> List[int]:
        return sorted({x for t in self.trees for x in t.nodes})

    def validate(self, world: Optional[int] = None) -> None:
        for i, t in enumerate(self.trees):
            if len(set(t.nodes)) != len(t.nodes):
                raise StrategyError(f"tree {i}: duplicate rank")
            if world is not None:
                missing = set(range(world)) - set(t.nodes)
                if missing:
                    raise StrategyError(f"tree {i}: ranks {sorted(missing)} missing")


def _add_subtree(x: xmlio.Node, parent: int, t: Tree) -> None:
    if "id" not in x.attrs:
        raise StrategyError(f"<{x.tag}> without id")
    try:
        rank = int(x.attrs["id"])
    except ValueError as e:
        raise StrategyError(f"bad rank id {x.attrs['id']!r}") from e
    if rank < 0:
        raise StrategyError("negative rank id")
    if rank in t.parent or rank == t.root:
        raise StrategyError(f"rank {rank} appears twice in one tree")
    if parent < 0:
        t.root = rank
    else:
        t.parent[rank] = parent
        t.children.setdefault(parent, []).append(rank)
    t.nodes.append(rank)
    t.ip[rank] = x.attrs.get("ip", "")
    for ch in x.children:
        if ch.tag == "gpu":
            _add_subtree(ch, rank, t)


# ---- generators (shapes the synthesizer can emit on a uniform NVSwitch) --------------------
def _mk(order: Sequence[int], parent_of: Callable[[int], int], ips: Optional[Sequence[str]]) -> Tree:
    t = Tree(root=order[0])
    for i, r in enumerate(order):
        t.nodes.append(r)
        t.ip[r] = ips[r] if ips else "127.0.0.1"
        if i:
            p = order[parent_of(i)]
            t.parent[r] = p
            t.children.setdefault(p, []).append(r)
    # nodes must be DFS pre-order for contraction semantics
    out: List[int] = []

    def dfs(x: int) -> None:
        out.append(x)
        for c in t.kids(x):
            dfs(c)
    dfs(t.root)
    t.nodes = out
    return t


def chain_tree(order: Sequence[int], ips=None) -> Tree:
    return _mk(order, lambda i: i - 1, ips)


def binary_tree(order: Sequence[int], ips=None) -> Tree:
    return _mk(order, lambda i: (i - 1) // 2, ips)


def star_tree(order: Sequence[int], ips=None) -> Tree:
    return _mk(order, lambda i: 0, ips)


def kary_tree(order: Sequence[int], k: int, ips=None) -> Tree:
    return _mk(order, lambda i: (i - 1) // k, ips)


def rotated(world: int, shift: int) -> List[int]:
    return [(r + shift) % world for r in range(world)]


def make_strategy(world: int, degree: int, shape: str = "binary", ips=None) -> Strategy:
    """``degree`` rotated trees of the given shape over ranks 0..world-1 (every rank is a root of
    at most one tree, so root work is spread like the reference's rotated binary trees,
    /root/reference/gurobi/trees.py:133-139)."""
    degree = max(1, min(degree, world))
    fn = {"binary": binary_tree, "chain": chain_tree, "star": star_tree}.get(shape)
    if fn is None:
        raise StrategyError(f"unknown shape {shape!r}")
    step = max(1, world // degree)
    return Strategy([fn(rotated(world, i * step), ips) for i in range(degree)], {"shape": shape})
