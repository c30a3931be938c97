"""Schedule compiler: strategy trees + tensor size -> flat list of work items.

Mirrors the native plan (csrc/comm_context.cu tree_collective / kernels_tree.cuh): the tensor is
cut into 16-byte packs of the wire dtype, split into one contiguous slice per tree and each slice
into chunks of ``chunk_bytes``. Unlike the reference (integer division drops ``size % numTrans``
and ``tranSize % chunkFloatNum`` elements, /root/reference/csrc/allreduce.cu:536-539) every element
belongs to exactly one (tree, chunk) item.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List


@dataclass(frozen=True)
class WorkItem:
    tree: int
    chunk: int
    start: int      # first element
    length: int     # number of elements (clipped to the tensor)


def wire_epp(wire_itemsize: int) -> int:
    """Elements per 16-byte pack."""
    return 16 // wire_itemsize


def slice_bounds(count: int, n_trees: int, wire_itemsize: int = 4) -> List[int]:
    """Element offsets [b0..b_n]: tree t owns [b_t, b_{t+1})."""
    epp = wire_epp(wire_itemsize)
    npacks = -(-count // epp)
    per = -(-npacks // n_trees) if n_trees else 0
    return [min(min(t * per, npacks) * epp, count) for t in range(n_trees + 1)]


def work_items(count: int, n_trees: int, chunk_bytes: int, wire_itemsize: int = 4) -> List[WorkItem]:
    """Items in execution order: chunk-major, trees interleaved (all trees progress together)."""
    epp = wire_epp(wire_itemsize)
    chunk_elems = max(16, chunk_bytes) // 16 * epp
    b = slice_bounds(count, n_trees, wire_itemsize)
    per_tree = [-(-(b[t + 1] - b[t]) // chunk_elems) for t in range(n_trees)]
    items: List[WorkItem] = []
    for k in range(max(per_tree, default=0)):
        for t in range(n_trees):
            if k < per_tree[t]:
                s = b[t] + k * chunk_elems
                items.append(WorkItem(t, k, s, min(chunk_elems, b[t + 1] - s)))
    return items


def default_chunk_bytes(total_bytes: int) -> int:
    """The reference hook's rule (/root/reference/commu.py:399-403): 4 MiB chunks above 10 MiB,
    else a quarter of the bucket — rounded to whole packs."""
    if total_bytes > 10 * 1024 * 1024:
        return 4 * 1024 * 1024
    return max(16, (total_bytes // 4) // 16 * 16)
