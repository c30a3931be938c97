#!/usr/bin/env python
"""DDP gradient-bucket sizes of the workload models (float counts, bytes, the hook's chunk size) — the counterpart of the
reference's hand-kept log/model_bucket_info.txt. Uses torch's own bucket assignment (first bucket 1 MB, then the cap),
in reverse parameter order like DDP does before its first rebuild. CPU only:

    python tools/bucket_info.py [--cap_mb 25] > log/model_bucket_info.txt
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from adapcc_b200.strategy.schedule import default_chunk_bytes  # noqa: E402


def buckets_of(model, cap_mb):
    params = [p for p in model.parameters() if p.requires_grad][::-1]
    idx, _ = torch._C._distributed_c10d._compute_bucket_assignment_by_size(params, [1 << 20, cap_mb << 20])
    return [sum(params[i].numel() for i in b) for b in idx]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cap_mb", type=int, default=25)
    a = ap.parse_args()
    import torchvision.models as tv

    from adapcc_b200.models.gpt2 import GPT2Config, GPT2DoubleHeads
    from adapcc_b200.models.vit import ViT, ViTConfig
    models = {"resnet18": lambda: tv.resnet18(), "vgg16": lambda: tv.vgg16(),
              "gpt2-small double heads": lambda: GPT2DoubleHeads(GPT2Config()),
              "vit (reference shape: dim 1024, depth 6)": lambda: ViT(ViTConfig.reference())}
    print(f"# DDP buckets (bucket_cap_mb={a.cap_mb}, fp32 gradients; chunk = the hook's pipelining granularity)")
    for name, make in models.items():
        with torch.device("meta"):
            m = make()
        sizes = buckets_of(m, a.cap_mb)
        total = sum(sizes)
        print(f"\n{name}: {total} parameters, {len(sizes)} buckets")
        for i, n in enumerate(sizes):
            print(f"  bucket {i}: {n} floats, {4 * n} bytes, chunk {default_chunk_bytes(4 * n)} bytes")


if __name__ == "__main__":
    main()
