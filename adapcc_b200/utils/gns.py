"""Gradient noise scale probes — the reference's units-test/get_gns.py (``GNS.compute_sample_grads``,
``compute_gns`` vector / split / whole variants, /root/reference/units-test/get_gns.py:1-108) and the
probe hooked into its ImageNet trainer. B_simple = tr(Sigma) / |G|^2 estimated from gradient norms at
two batch sizes (McCandlish et al.): with per-rank gradients g_i (batch b) and their mean G (batch
n*b):   |G|^2_est = (n*b*|G|^2 - b*mean|g_i|^2) / (n*b - b),   S_est = (mean|g_i|^2 - |G|^2) / (1/b - 1/(n*b)).

In a data-parallel job the small-batch norms are free: every rank has its local gradient before the
all-reduce and the averaged one after it, so ``GNSProbe`` needs two extra scalars per step."""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch


def grad_sq_norm(params: Iterable[torch.Tensor]) -> torch.Tensor:
    total = None
    for p in params:
        g = p.grad if isinstance(p, torch.nn.Parameter) or hasattr(p, "grad") and p.grad is not None else p
        if g is None:
            continue
        s = g.detach().float().pow(2).sum()
        total = s if total is None else total + s
    return total if total is not None else torch.zeros(())


def compute_gns(small_sq_norms: torch.Tensor, big_sq_norm: torch.Tensor, b_small: int, b_big: int):
    """-> (gns, |G|^2 estimate, trace(Sigma) estimate). ``small_sq_norms``: |g_i|^2 of every small batch."""
    gs = small_sq_norms.float().mean()
    gb = big_sq_norm.float()
    g2 = (b_big * gb - b_small * gs) / (b_big - b_small)
    s = (gs - gb) / (1.0 / b_small - 1.0 / b_big)
    return s / g2.clamp(min=1e-30), g2, s


class GNSProbe:
    """EMA-smoothed GNS from (local grad norm^2 before all-reduce, averaged grad norm^2 after)."""

    def __init__(self, world_size: int, local_batch: int, beta: float = 0.95):
        self.n, self.b, self.beta = world_size, local_batch, beta
        self.ema_s: Optional[float] = None
        self.ema_g2: Optional[float] = None
        self.history: List[float] = []

    def update(self, mean_local_sq: float, global_sq: float) -> Optional[float]:
        if self.n < 2:
            return None
        _, g2, s = compute_gns(torch.tensor([mean_local_sq]), torch.tensor(global_sq), self.b, self.b * self.n)
        self.ema_s = float(s) if self.ema_s is None else self.beta * self.ema_s + (1 - self.beta) * float(s)
        self.ema_g2 = float(g2) if self.ema_g2 is None else self.beta * self.ema_g2 + (1 - self.beta) * float(g2)
        gns = self.ema_s / max(self.ema_g2, 1e-30)
        self.history.append(gns)
        return gns


def compute_sample_grads(model: torch.nn.Module, loss_fn, inputs: torch.Tensor, targets: torch.Tensor):
    """Per-sample gradients with torch.func (the reference loops sample by sample)."""
    from torch.func import functional_call, grad, vmap

    params = {k: v.detach() for k, v in model.named_parameters()}
    buffers = {k: v.detach() for k, v in model.named_buffers()}

    def one(p, b, x, y):
        out = functional_call(model, (p, b), (x.unsqueeze(0),))
        return loss_fn(out, y.unsqueeze(0))

    return vmap(grad(one), in_dims=(None, None, 0, 0))(params, buffers, inputs, targets)
