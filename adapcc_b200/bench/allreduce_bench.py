"""Bus-bandwidth sweep of our collectives next to NCCL — BASELINE.json's first headline metric
("allreduce bus GB/s vs message size, 1 KB - 1 GB, 2/4/8 GPUs, per-strategy XML").

Same definitions as the nccl-tests copy the reference ships (/root/reference/nccl-perf/benchmark/
PERFORMANCE.md:33-63,134-142; src/all_reduce.cu:48-54): algbw = bytes / time, busbw = algbw *
2(n-1)/n for all-reduce (factor 1 for reduce / broadcast). Timing: every variant is captured as a
CUDA graph of ``iters`` back-to-back calls and replayed, timed with CUDA events on the launching
stream, best of 3, MAX over ranks — so launch overhead is excluded equally for NCCL and for us.
Each message size uses buffers larger than the previous call's (and for >= 128 MB exceeds L2).

    torchrun --nproc-per-node 8 -m adapcc_b200.bench.allreduce_bench --out gpurun_out/sweep8.json
"""
from __future__ import annotations

import argparse
import json
import os

import torch
import torch.distributed as dist

from ..constants import ALLREDUCE
from ..runtime.native import NativeComm
from ..runtime.rendezvous import unique_name
from ..strategy import make_strategy


def timeit(fn, iters, dev, side, graph=True):
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
        side.synchronize()
        g = None
        if graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for _ in range(iters):
                    fn()
            g.replay()
            side.synchronize()
        best = 1e9
        for _ in range(3):
            dist.barrier()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(side)
            if g is not None:
                g.replay()
            else:
                for _ in range(iters):
                    fn()
            e.record(side)
            side.synchronize()
            best = min(best, s.elapsed_time(e) / iters)
    t = torch.tensor([best], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min_log2", type=int, default=10)
    ap.add_argument("--max_log2", type=int, default=30)
    ap.add_argument("--step", type=int, default=1)
    ap.add_argument("--dtype", default="float32", choices=["float32", "bfloat16"])
    ap.add_argument("--blocks", type=int, default=0, help="CTAs per collective (0 = library default)")
    ap.add_argument("--out", default="")
    ap.add_argument("--no_tree", action="store_true")
    ap.add_argument("--pipe_sweep", action="store_true", help="stager/link CTA split sensitivity of the staged path")
    a = ap.parse_args()

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    dtype = getattr(torch, a.dtype)
    esize = torch.empty((), dtype=dtype).element_size()
    max_bytes = 1 << a.max_log2
    comm = NativeComm(unique_name("sweep"), rank, world, local, staging_bytes=min(max_bytes, 1 << 30),
                      heap_bytes=max_bytes + (4 << 20))
    if a.blocks:
        comm.set_tunable("max_blocks", a.blocks)
        comm.set_tunable("tree_blocks", a.blocks)
    comm.load_strategy(make_strategy(world, min(4, world), "binary").to_xml())
    side = torch.cuda.Stream()
    factor = 2 * (world - 1) / world
    rows = []
    algos = ["one_shot", "two_shot"] + (["nvls"] if comm.multicast else [])
    if rank == 0:
        print(f"# world={world} dtype={a.dtype} symm={comm.symm_backend} multicast={comm.multicast} "
              f"busbw = algbw*{factor:.3f} (GB/s), time in us, device-timed (graph replay), max over ranks", flush=True)
    for p in range(a.min_log2, a.max_log2 + 1, a.step):
        nbytes = 1 << p
        n = nbytes // esize
        x = torch.randn(n, device=dev).to(dtype)
        comm.heap_reset()
        hz = comm.symm_empty(n, dtype)
        iters = 40 if nbytes <= (1 << 22) else (10 if nbytes <= (1 << 26) else 3)
        row = {"bytes": nbytes}
        row["nccl"] = timeit(lambda: dist.all_reduce(x), iters, dev, side)
        for algo in algos:
            if algo == "one_shot" and nbytes > (8 << 20):
                continue
            row[algo] = timeit(lambda: comm.all_reduce(x, algo=algo), iters, dev, side)
            if algo != "one_shot":
                row[algo + "_zc"] = timeit(lambda: comm.all_reduce(hz, algo=algo), iters, dev, side)
        row["auto"] = timeit(lambda: comm.all_reduce(x, algo="auto"), iters, dev, side)
        row["auto_zc"] = timeit(lambda: comm.all_reduce(hz, algo="auto"), iters, dev, side)
        if dtype == torch.float32:
            row["auto_bf16wire"] = timeit(lambda: comm.all_reduce(x, algo="auto", wire="bfloat16"), iters, dev, side)
        if not a.no_tree and nbytes >= (1 << 16):
            row["tree"] = timeit(lambda: comm.tree_collective(ALLREDUCE, x, chunk_bytes=4 << 20), iters, dev, side)
        comm.check()
        ours = min(v for k, v in row.items() if k not in ("bytes", "nccl"))
        row["best_vs_nccl"] = row["nccl"] / ours
        rows.append(row)
        if rank == 0:
            print("%11d B  " % nbytes + "  ".join(
                f"{k}={v * 1e6:8.1f}us/{nbytes * factor / v / 1e9:6.1f}" for k, v in row.items()
                if k not in ("bytes", "best_vs_nccl")) + f"  best/nccl={row['best_vs_nccl']:.2f}x", flush=True)
        del x
    if a.pipe_sweep:
        for nbytes in (1 << 28, 1 << 30):
            if nbytes > max_bytes:
                continue
            n = nbytes // esize
            x = torch.randn(n, device=dev).to(dtype)
            for st, ln in ((0, 0), (32, 32), (48, 48), (64, 64), (40, 72), (72, 40)):
                if st == 0:
                    comm.set_tunable("pipe_min_bytes", 0)
                else:
                    comm.set_tunable("pipe_min_bytes", 32 << 20)
                    comm.set_tunable("pipe_stagers", st)
                    comm.set_tunable("pipe_links", ln)
                t = timeit(lambda: comm.all_reduce(x, algo="auto"), 3, dev, side)
                comm.check()
                rows.append({"bytes": nbytes, "pipe_stagers": st, "pipe_links": ln, "auto": t})
                if rank == 0:
                    print(f"[pipe] {nbytes} B stagers={st} links={ln}: {t * 1e6:9.1f} us {nbytes * factor / t / 1e9:7.1f} GB/s", flush=True)
            del x
        comm.set_tunable("pipe_min_bytes", 32 << 20)
        comm.set_tunable("pipe_stagers", 48)
        comm.set_tunable("pipe_links", 48)
    if rank == 0 and a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        with open(a.out, "w") as f:
            json.dump({"world": world, "dtype": a.dtype, "busbw_factor": factor, "multicast": comm.multicast,
                       "timing": "cuda-graph replay, cuda events, best of 3, max over ranks", "rows": rows}, f, indent=1)
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
