"""``python -m adapcc_b200.doctor`` — environment and connectivity check before a job is launched.

The reference's pre-flight checks are ``units-test/check_mpi_connect.py`` (``mpirun … echo HELLO`` on every host),
``units-test/check-p2p`` (a 2-rank CUDA-aware MPI ping-pong) and the RPC latency dumps under ``proto/latency_*.txt``.
Here one command reports: toolchain and library state, GPUs / peer access / NVLink / multicast as the native detector sees
them, a multi-process rendezvous on this host (gloo or NCCL), and the coordinator's RPC round trip (same statistics as the
reference's latency files). Exit code 0 = nothing that would stop a job was found.

    python -m adapcc_b200.doctor [--ranks 2] [--rpc 1000] [--json]
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import time


def _toolchain() -> dict:
    import torch

    out = {"python": sys.version.split()[0], "torch": torch.__version__, "torch_cuda": torch.version.cuda,
           "nvcc": shutil.which("nvcc") or ("/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else None)}
    try:
        out["nccl"] = ".".join(map(str, torch.cuda.nccl.version()))
    except Exception:  # noqa: BLE001
        out["nccl"] = None
    for mod in ("grpc", "scipy", "numpy", "torchvision"):
        try:
            out[mod] = __import__(mod).__version__
        except Exception as e:  # noqa: BLE001
            out[mod] = f"missing ({type(e).__name__})"
    return out


def _library() -> dict:
    from .build import LIB
    from .runtime.native import load_library

    out = {"path": str(LIB), "built": LIB.exists()}
    if not LIB.exists():
        out["hint"] = "python -m adapcc_b200.build"
        return out
    try:
        lib = load_library(build_if_missing=False)
        want = ("initThreads", "exitThreads", "allreduce", "reduce", "boardcast", "updateActive", "adapcc_allreduce",
                "adapcc_tree_collective", "adapcc_detect_topology", "adapcc_profile_links", "adapcc_gemm_pp")
        out["missing_symbols"] = [s for s in want if not hasattr(lib, s)]
    except OSError as e:
        out["load_error"] = str(e)
    return out


def _gpus() -> dict:
    import torch

    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    out = {"count": n, "devices": []}
    for d in range(n):
        p = torch.cuda.get_device_properties(d)
        out["devices"].append({"index": d, "name": p.name, "sm": f"{p.major}{p.minor}", "sms": p.multi_processor_count,
                               "memory_gb": round(p.total_memory / 2 ** 30, 1)})
    if n > 1:
        out["peer_access"] = [[bool(a == b or torch.cuda.can_device_access_peer(a, b)) for b in range(n)] for a in range(n)]
        out["all_peers"] = all(all(r) for r in out["peer_access"])
    if n:
        try:                                   # what the native detector sees (NVLink counts, NVSwitch, multicast)
            import ctypes
            import re

            from .runtime.native import load_library

            lib = load_library(build_if_missing=False)
            buf = ctypes.create_string_buffer(1 << 20)
            if lib.adapcc_detect_topology(0, buf, len(buf)) >= 0:
                xml = buf.value.decode()
                out["nvlinks"] = [int(x) for x in re.findall(r'nvlinks="(\d+)"', xml)]
                out["nvswitch_links"] = [int(x) for x in re.findall(r'nvswitch_links="(\d+)"', xml)]
                out["multicast"] = [int(x) for x in re.findall(r'multicast="(\d+)"', xml)]
        except Exception as e:  # noqa: BLE001
            out["detect_error"] = str(e)
    return out


_HELLO = r"""
import os, sys, time, torch, torch.distributed as dist
backend = sys.argv[1]
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
if backend == "nccl":
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group(backend)
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"])) if backend == "nccl" else torch.device("cpu")
t = torch.ones(1024, device=dev) * (rank + 1)
t0 = time.time(); dist.all_reduce(t)
if backend == "nccl": torch.cuda.synchronize()
ok = bool((t == world * (world + 1) / 2).all())
if rank == 0: print("HELLO", world, ok, round((time.time() - t0) * 1e3, 2), flush=True)
dist.destroy_process_group()
"""


def _rendezvous(ranks: int, backend: str) -> dict:
    import socket
    import tempfile

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write(_HELLO)
        script = f.name
    try:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), script, backend]
        t0 = time.time()
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("HELLO")]
        ok = r.returncode == 0 and bool(line) and line[0].split()[2] == "True"
        return {"backend": backend, "ranks": ranks, "ok": ok, "seconds": round(time.time() - t0, 1),
                "first_allreduce_ms": float(line[0].split()[3]) if line else None,
                "stderr_tail": "" if ok else r.stderr[-600:]}
    finally:
        os.unlink(script)


def _rpc(n: int) -> dict:
    """Round trip of ``hook_fetch`` through a real gRPC server on loopback (world size 1: the decision is immediate, so
    this is protocol + transport cost — what /root/reference/proto/latency_0.0.txt records: median 0.99 ms)."""
    from .coord.client import Hooker
    from .coord.server import Coordinator, make_server

    c = Coordinator("127.0.0.1", 0, 1)
    srv = make_server(c)
    srv.start()
    try:
        h = Hooker("127.0.0.1", c.port, timeout=10)
        for s in range(20):
            h.send_ready_request(s, 0)
        lat = []
        for s in range(20, 20 + n):
            t0 = time.perf_counter()
            h.send_ready_request(s, 0)
            lat.append((time.perf_counter() - t0) * 1e3)
        h.close()
    finally:
        srv.stop(0)
    lat.sort()
    return {"calls": n, "mean_ms": round(statistics.mean(lat), 3), "median_ms": round(statistics.median(lat), 3),
            "p95_ms": round(lat[int(0.95 * (len(lat) - 1))], 3), "max_ms": round(lat[-1], 3),
            "reference_median_ms": 0.99}


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--ranks", type=int, default=2, help="processes of the rendezvous check (0: skip)")
    ap.add_argument("--rpc", type=int, default=200, help="coordinator round trips to time (0: skip)")
    ap.add_argument("--json", action="store_true")
    a = ap.parse_args(argv)
    rep = {"toolchain": _toolchain(), "library": _library(), "gpus": _gpus()}
    if a.ranks > 0:
        n_gpu = rep["gpus"]["count"]
        backend = "nccl" if n_gpu >= a.ranks else "gloo"
        rep["rendezvous"] = _rendezvous(a.ranks, backend)
    if a.rpc > 0:
        rep["coordinator_rpc"] = _rpc(a.rpc)
    problems = []
    if not rep["library"].get("built"):
        problems.append("native library not built: python -m adapcc_b200.build")
    if rep["library"].get("missing_symbols") or rep["library"].get("load_error"):
        problems.append(f"native library incomplete: {rep['library']}")
    if rep["gpus"]["count"] > 1 and not rep["gpus"].get("all_peers", True):
        problems.append("not every GPU pair has peer access: the in-kernel NVLink path needs it")
    if "rendezvous" in rep and not rep["rendezvous"]["ok"]:
        problems.append("multi-process rendezvous on 127.0.0.1 failed")
    if str(rep["toolchain"].get("grpc", "")).startswith("missing"):
        problems.append("grpcio missing: relay control / fault detection need the coordinator")
    rep["problems"] = problems
    if a.json:
        print(json.dumps(rep, indent=1))
    else:
        t = rep["toolchain"]
        print(f"python {t['python']}  torch {t['torch']} (CUDA {t['torch_cuda']}, NCCL {t['nccl']})  nvcc {t['nvcc']}  "
              f"grpc {t['grpc']}  scipy {t['scipy']}")
        lib = rep["library"]
        print(f"native library: {lib['path']}  built={lib['built']}  missing symbols={lib.get('missing_symbols', '?')}")
        g = rep["gpus"]
        print(f"GPUs: {g['count']}" + "".join(f"\n  [{d['index']}] {d['name']} sm_{d['sm']} {d['sms']} SMs {d['memory_gb']} GB"
                                             for d in g["devices"]))
        if g["count"] > 1:
            print(f"  peer access between all pairs: {g.get('all_peers')}  NVLinks per GPU: {g.get('nvlinks')}  "
                  f"to NVSwitch: {g.get('nvswitch_links')}  multicast: {g.get('multicast')}")
        if "rendezvous" in rep:
            r = rep["rendezvous"]
            print(f"rendezvous: {r['ranks']} x {r['backend']} on 127.0.0.1: {'ok' if r['ok'] else 'FAILED'} "
                  f"({r['seconds']} s, first all-reduce {r['first_allreduce_ms']} ms){' ' + r['stderr_tail'] if not r['ok'] else ''}")
        if "coordinator_rpc" in rep:
            c = rep["coordinator_rpc"]
            print(f"coordinator RPC round trip over loopback gRPC ({c['calls']} calls): mean {c['mean_ms']} ms, median "
                  f"{c['median_ms']} ms, p95 {c['p95_ms']} ms, max {c['max_ms']} ms (reference: median {c['reference_median_ms']} ms)")
        print("problems: " + ("none" if not problems else "; ".join(problems)))
    return 1 if problems else 0


if __name__ == "__main__":
    raise SystemExit(main())
