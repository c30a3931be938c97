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


def test_native_entry_points_fail_loudly_without_a_gpu():
    """On a box without a GPU the native runtime reports an error (message available) instead of crashing
    or silently falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from adapcc_b200.runtime.native import load_library
    lib = load_library()
    lib.adapcc_last_error.restype = ctypes.c_char_p
    lib.adapcc_detect_topology.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
    buf = ctypes.create_string_buffer(4096)
    assert lib.adapcc_detect_topology(0, buf, 4096) < 0
    assert b"no CUDA device" in lib.adapcc_last_error()
    lib.adapcc_ctx_create.restype = ctypes.c_void_p
    lib.adapcc_ctx_create.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_ulonglong,
                                      ctypes.c_ulonglong]
    assert lib.adapcc_ctx_create(b"cpu-only", 0, 1, 0, 1 << 20, 0) is None
    assert b"cudaSetDevice" in lib.adapcc_last_error()
