"""Relay control: what a rank does in one tree when only a subset of GPUs is active.

Pure-Python mirror of csrc/schedule.cpp (cross-checked in tests) and behavioural parity with
/root/reference/csrc/control.cu:7-101:

* ``has_recv``   — some child's subtree contains an active rank,
* ``has_local``  — this rank is active (contributes its own data),
* ``has_kernel`` — a reduction runs here (data arrives AND it is not a single flow merely passing
  through an inactive rank),
* ``has_send``   — (active or receiving) and not the root.

Golden rows from the reference's logs (tree 0<-1<-{2,3}):
  all active  -> r0 (1,1,1,0) r1 (1,1,1,1) r2 (0,1,0,1) r3 (0,1,0,1)   /root/reference/log/primitive:139-146
  active{0,2} -> r0 (1,1,1,0) r1 (1,0,0,1) r2 (0,1,0,1) r3 (0,0,0,0)   /root/reference/log/training:150-159
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Iterable, List, Set

from ..constants import (ALLREDUCE, BOARDCAST, REDUCE, RELAY_BYPASS, RELAY_FORWARD, TR_HAS_LOCAL,
                         TR_IN_BCAST, TR_IN_REDUCE, TR_PARENT_IS_ROOT, TR_PUBLISH, TR_WANT_RESULT)
from .trees import Strategy, Tree


@dataclass
class RelayControl:
    """The reference's relay truth table for one rank in one tree: does it receive, contribute its own data, launch a
    reduce kernel, send (``hasRecv / hasLocal / hasKernel / hasSend``, /root/reference/csrc/control.cu:72-101)."""

    has_recv: bool = False
    has_local: bool = False
    has_kernel: bool = False
    has_send: bool = False
    active_recvs: List[int] = field(default_factory=list)

    def as_tuple(self):
        return (int(self.has_recv), int(self.has_local), int(self.has_kernel), int(self.has_send))


def subtree_active(tree: Tree, x: int, active: Set[int]) -> bool:
    if x in active:
        return True
    return any(subtree_active(tree, c, active) for c in tree.kids(x))


def relay_control(tree: Tree, rank: int, active: Iterable[int]) -> RelayControl:
    act = set(active)
    rc = RelayControl()
    rc.active_recvs = [c for c in tree.kids(rank) if subtree_active(tree, c, act)]
    rc.has_recv = bool(rc.active_recvs)
    rc.has_local = rank in act
    rc.has_kernel = rc.has_recv and not (len(rc.active_recvs) == 1 and not rc.has_local)
    rc.has_send = (rc.has_local or rc.has_recv) and rank != tree.root and rank in tree.nodes
    return rc


@dataclass
class TreeRole:
    """What a rank does in one tree for one op after relay control: effective parent / children over the active subset
    (forward: inactive ranks stay on the path; bypass: they are contracted out) and role flags — mirrored natively
    by csrc/schedule.cpp::tree_role."""

    parent: int = -1
    children: List[int] = field(default_factory=list)
    flags: int = 0

    def any(self) -> bool:
        return self.flags != 0

    def has(self, f: int) -> bool:
        return bool(self.flags & f)


def tree_role(tree: Tree, rank: int, active: Iterable[int], prim: int = ALLREDUCE,
              relay_mode: int = RELAY_FORWARD) -> TreeRole:
    """Role of ``rank`` in ``tree`` for one op. In RELAY_BYPASS mode inactive ranks are contracted
    out first: on a uniform NVSwitch nobody needs a forwarder, the parent pulls straight from the
    nearest active descendant."""
    act = set(active)
    T = tree
    if relay_mode == RELAY_BYPASS:
        T = tree.contract(lambda x: x in act or (prim == BOARDCAST and x == tree.root))
    role = TreeRole()
    if rank not in T.nodes:
        return role
    is_root = rank == T.root
    local = rank in act
    recvs = [c for c in T.kids(rank) if subtree_active(T, c, act)]
    role.parent = -1 if is_root else T.parent.get(rank, -1)
    pir = TR_PARENT_IS_ROOT if (role.parent >= 0 and role.parent == T.root) else 0
    if prim in (ALLREDUCE, REDUCE):
        if not (local or recvs):
            return TreeRole()
        role.flags |= TR_IN_REDUCE
        if local:
            role.flags |= TR_HAS_LOCAL
        role.children = recvs
        if prim == ALLREDUCE:
            if not is_root:
                role.flags |= TR_IN_BCAST | pir
            if local:
                role.flags |= TR_WANT_RESULT
            if recvs:
                role.flags |= TR_PUBLISH
        elif is_root:
            role.flags |= TR_WANT_RESULT
    elif prim == BOARDCAST:
        if is_root:
            role.flags |= TR_PUBLISH
        else:
            if not (local or recvs):
                return TreeRole()
            role.flags |= TR_IN_BCAST | pir
            if local:
                role.flags |= TR_WANT_RESULT
            if recvs:
                role.flags |= TR_PUBLISH
    else:
        raise ValueError(f"primitive {prim} has no tree schedule")
    return role


def participants(strategy: Strategy, world: int, active: Iterable[int], prim: int = ALLREDUCE,
                 relay_mode: int = RELAY_FORWARD) -> List[int]:
    """Ranks that hold any role in any tree (they synchronise at the end of the op)."""
    act = set(active)
    return [r for r in range(world)
            if any(tree_role(t, r, act, prim, relay_mode).any() for t in strategy.trees)]
