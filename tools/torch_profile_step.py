#!/usr/bin/env python
"""Kernel-level time table of one GPT-2 training step (eager engine, 1 GPU) from torch.profiler.

ncu serialises kernels and replays them cache-cold (and costs GPU-minutes); CUPTI activity records
give the in-situ durations of every kernel of a real step in a few seconds:

    python tools/torch_profile_step.py --out gpurun_out/torch_profile.md [--steps 3]

The table groups kernels by (shortened) name: launches per step, total us per step, share.
"""
from __future__ import annotations

import argparse
import os
import re
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def short(name: str) -> str:
    name = re.sub(r"<.*", "", name)
    name = re.sub(r"^void\s+", "", name)
    name = name.replace("at::native::", "native::").replace("(anonymous namespace)::", "")
    return name[:90]


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/torch_profile.md")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--candidates", type=int, default=2)
    ap.add_argument("--seq", type=int, default=1024)
    ap.add_argument("--lm_rows", default="all", choices=["all", "scored"])
    a = ap.parse_args()

    import torch
    from torch.profiler import ProfilerActivity, profile

    from adapcc_b200.models.gpt2 import GPT2Config, GPT2DoubleHeads, lm_rows_needed, synthetic_batch
    from adapcc_b200.parallel.engine import FlatDataParallel

    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    cfg = GPT2Config()
    torch.manual_seed(0)
    model = GPT2DoubleHeads(cfg).to(dev)
    batch = synthetic_batch(a.batch, a.candidates, a.seq, cfg.vocab_size, device=dev)
    if a.lm_rows == "scored":
        model.lm_row_capacity = lm_rows_needed(batch["lm_labels"].cpu())
    eng = FlatDataParallel(model, None, world_size=1, lr=1e-4, max_norm=1.0)
    for _ in range(3):
        eng.step(batch)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(a.steps):
            eng.step(batch)
        torch.cuda.synchronize()
    rows = {}
    total = 0.0
    for ev in prof.events():
        if ev.device_type is not None and str(ev.device_type).endswith("CUDA") and ev.device_time_total > 0:
            k = short(ev.name)
            n, t = rows.get(k, (0, 0.0))
            rows[k] = (n + 1, t + ev.device_time_total)
            total += ev.device_time_total
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        f.write(f"# GPT-2 small step, eager engine, torch.profiler (CUPTI) kernel durations, {a.steps} steps, "
                f"lm_rows={a.lm_rows}, fuse_add_ln={model.fuse_add_ln}\n\n")
        f.write(f"kernel time per step: {total / a.steps / 1e3:.3f} ms "
                f"({sum(n for n, _ in rows.values()) / a.steps:.0f} launches)\n\n")
        f.write("| kernel | launches/step | us/step | share |\n|---|---|---|---|\n")
        for k, (n, t) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:60]:
            f.write(f"| {k} | {n / a.steps:.1f} | {t / a.steps:.1f} | {100 * t / total:.1f}% |\n")
    print(open(a.out).read()[:3000])
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
