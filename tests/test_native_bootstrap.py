"""The native Unix-socket rendezvous (csrc/bootstrap.cpp) on CPU: N processes build the mesh, allgather,
pass file descriptors with SCM_RIGHTS (the mechanism that carries CUDA VMM handles between the per-GPU
processes) and meet in a barrier — no GPU involved."""
import ctypes
import multiprocessing as mp
import os

import pytest


def _worker(name, rank, world, q):
    try:
        from adapcc_b200.runtime.native import load_library
        lib = load_library()
        lib.adapcc_bootstrap_selftest.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        lib.adapcc_last_error.restype = ctypes.c_char_p
        rc = lib.adapcc_bootstrap_selftest(name.encode(), rank, world, 20000)
        q.put((rank, rc, lib.adapcc_last_error().decode() if rc else ""))
    except Exception as e:                                     # pragma: no cover
        q.put((rank, -99, repr(e)))


@pytest.mark.parametrize("world", [2, 5])
def test_bootstrap_mesh_allgather_fd_passing_barrier(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    name = f"adapcc-selftest-{os.getpid()}-{world}"
    procs = [ctx.Process(target=_worker, args=(name, r, world, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert sorted(r for r, _, _ in results) == list(range(world))
    assert all(rc == 0 for _, rc, _ in results), results
