"""MoE transformer MLP workload (BASELINE.json config 5: "MoE DDP with relay control").

Reference: ``fmoe.FMoETransformerMLP(num_expert=10, d_model=1024, d_hidden=4096, top_k=1)`` wrapped
in ``DistributedGroupedDataParallel`` with local experts only (/root/reference/models/moe/train_moe.py:37-54).
Here: top-k softmax gate, per-expert 2-layer MLP as ONE batched GEMM pair over a capacity-padded
[E, rows, d] buffer, and two exchange modes:

* ``exchange=None``  — experts local to the rank (the reference's configuration); tokens are
  scattered into the expert-major buffer by the same assign/push kernels with world = 1;
* ``exchange=ExpertExchange`` — expert parallel: experts sharded over the ranks, tokens exchanged
  by in-kernel peer stores/loads over NVLink (adapcc_b200/parallel/expert_parallel.py).

The gate and any dense parameters are data-parallel (all-reduced by the DDP hook); expert
parameters are rank-local in expert-parallel mode (fastmoe's ``dp_comm="none"``).
"""
from __future__ import annotations

import math
import torch
import torch.nn as nn
import torch.nn.functional as F


class MoEMLP(nn.Module):
    """Top-k gated mixture-of-experts MLP (``num_expert`` × (d_model → d_hidden → d_model)), the layer of the
    reference's MoE workload (``fmoe.FMoETransformerMLP``, /root/reference/models/moe/train_moe.py:37-44). Experts
    local by default; with an ``ExpertExchange`` they are sharded over the ranks and tokens travel through the
    native dispatch / combine kernels."""

    def __init__(self, num_expert: int = 10, d_model: int = 1024, d_hidden: int = 4096, top_k: int = 1,
                 world_size: int = 1, capacity_factor: float = 2.0, exchange=None):
        super().__init__()
        self.num_expert, self.d_model, self.d_hidden, self.top_k = num_expert, d_model, d_hidden, top_k
        self.world_size, self.capacity_factor, self.exchange = world_size, capacity_factor, exchange
        self.total_expert = num_expert * world_size
        self.gate = nn.Linear(d_model, self.total_expert)
        self.w1 = nn.Parameter(torch.empty(num_expert, d_model, d_hidden))
        self.b1 = nn.Parameter(torch.zeros(num_expert, 1, d_hidden))
        self.w2 = nn.Parameter(torch.empty(num_expert, d_hidden, d_model))
        self.b2 = nn.Parameter(torch.zeros(num_expert, 1, d_model))
        for w in (self.w1, self.w2):
            nn.init.normal_(w, std=0.02)
        for p in (self.w1, self.b1, self.w2, self.b2):
            p.expert_parallel = True           # not all-reduced across data-parallel ranks in EP mode

    def capacity(self, n_tokens: int) -> int:
        per = n_tokens * self.top_k / self.total_expert
        return max(8, int(math.ceil(per * self.capacity_factor / 8) * 8))

    def experts(self, buf: torch.Tensor) -> torch.Tensor:
        """buf [E_local, rows, d] -> [E_local, rows, d]: two batched GEMMs (cuBLAS) + GELU."""
        h = torch.baddbmm(self.b1.to(buf.dtype), buf, self.w1.to(buf.dtype))
        return torch.baddbmm(self.b2.to(buf.dtype), F.gelu(h), self.w2.to(buf.dtype))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        shape = x.shape
        x = x.reshape(-1, self.d_model)
        T = x.shape[0]
        logits = self.gate(x).float()
        w, idx = torch.topk(F.softmax(logits, dim=-1), self.top_k, dim=-1)          # [T, k]
        if self.top_k > 1:
            w = w / w.sum(-1, keepdim=True)
        expert = idx.reshape(-1).to(torch.int32).contiguous()                        # [A]
        rows = x.repeat_interleave(self.top_k, dim=0) if self.top_k > 1 else x
        if self.exchange is not None and x.is_cuda:
            ex = self.exchange
            pos, _ = ex.assign(expert)
            buf = ex.dispatch(rows.to(torch.bfloat16), expert, pos)
            out = self.experts(buf)
            y = ex.combine(out, expert, pos, rows.shape[0]).to(x.dtype)
        else:
            y = self._local(rows, expert)
        y = (y.view(T, self.top_k, self.d_model) * w.unsqueeze(-1).to(y.dtype)).sum(1)
        return y.reshape(shape)

    def _local(self, rows: torch.Tensor, expert: torch.Tensor) -> torch.Tensor:
        """Reference-style local experts in plain PyTorch (CPU tests / oracle for the kernels):
        sort by expert, pad to capacity, batched GEMMs, un-sort. Same drop rule (capacity)."""
        A = rows.shape[0]
        cap = self.capacity(A // self.top_k)
        e = expert.long()
        order = torch.argsort(e, stable=True)
        se = e[order]
        start = torch.searchsorted(se, torch.arange(self.total_expert, device=e.device))
        pos = torch.arange(A, device=e.device) - start[se]
        keep = pos < cap
        buf = rows.new_zeros(self.total_expert, cap, self.d_model)
        buf[se[keep], pos[keep]] = rows[order][keep]
        out = self.experts(buf)
        y = rows.new_zeros(A, self.d_model)
        y[order[keep]] = out[se[keep], pos[keep]]
        return y
