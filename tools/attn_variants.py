"""Micro-experiment (1 GPU): fwd+bwd time of one GPT-2 block for different QKV layouts."""
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adapcc_b200.ops.layers import FusedLinear  # noqa: E402

dev = torch.device("cuda", 0)
B, T, D, H = 8, 1024, 768, 12


class Packed(nn.Module):
    def __init__(self):
        super().__init__()
        self.c_attn, self.c_proj = FusedLinear(D, 3 * D), FusedLinear(D, D)

    def forward(self, h):
        q, k, v = self.c_attn(h).view(B, T, 3, H, D // H).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v, is_causal=True)
        return self.c_proj(a.transpose(1, 2).reshape(B, T, D))


class Split(nn.Module):
    def __init__(self):
        super().__init__()
        self.q, self.k, self.v, self.c_proj = FusedLinear(D, D), FusedLinear(D, D), FusedLinear(D, D), FusedLinear(D, D)

    def forward(self, h):
        q = self.q(h).view(B, T, H, D // H).transpose(1, 2)
        k = self.k(h).view(B, T, H, D // H).transpose(1, 2)
        v = self.v(h).view(B, T, H, D // H).transpose(1, 2)
        a = F.scaled_dot_product_attention(q, k, v, is_causal=True)
        if not hasattr(self, "_printed"):
            self._printed = True
            print("   sdpa out strides", a.stride(), "transpose contiguous:", a.transpose(1, 2).is_contiguous())
        return self.c_proj(a.transpose(1, 2).reshape(B, T, D))


class SplitChunk(nn.Module):
    """one GEMM, split along the last dim (views), no permute"""

    def __init__(self):
        super().__init__()
        self.c_attn, self.c_proj = FusedLinear(D, 3 * D), FusedLinear(D, D)

    def forward(self, h):
        q, k, v = self.c_attn(h).split(D, dim=-1)
        q, k, v = (t.view(B, T, H, D // H).transpose(1, 2) for t in (q, k, v))
        a = F.scaled_dot_product_attention(q, k, v, is_causal=True)
        return self.c_proj(a.transpose(1, 2).reshape(B, T, D))


def bench(mod, name):
    mod = mod.to(dev).bfloat16()
    x = torch.randn(B, T, D, device=dev, dtype=torch.bfloat16, requires_grad=True)
    g = torch.randn(B, T, D, device=dev, dtype=torch.bfloat16)
    for _ in range(5):
        mod(x).backward(g)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        mod(x).backward(g)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:12s} {e0.elapsed_time(e1) / 50 * 1e3:8.1f} us per fwd+bwd")


for backend in ("default", "flash", "cudnn"):
    print("backend", backend)
    ctx = torch.nn.attention.sdpa_kernel({"flash": torch.nn.attention.SDPBackend.FLASH_ATTENTION,
                                          "cudnn": torch.nn.attention.SDPBackend.CUDNN_ATTENTION}[backend]) \
        if backend != "default" else __import__("contextlib").nullcontext()
    with ctx:
        try:
            bench(Packed(), "packed")
            bench(Split(), "split3")
            bench(SplitChunk(), "splitchunk")
        except Exception as e:  # noqa: BLE001
            print("  failed:", repr(e)[:200])
