"""tcgen05 GEMM with a fused bias + GELU epilogue (csrc/gemm_tcgen05.cu); numerically validated on B200, untuned.

``linear_act(x, weight, bias, act)`` computes ``act(x @ weight.T + bias)`` for bf16 CUDA tensors with the
accumulator in TMEM and the activation applied in the epilogue; ``linear_gelu`` is its autograd form (the
pre-activation is written by the same kernel for the backward). Opt-in: ``ADAPCC_TCGEN05_MLP=1`` makes the GPT-2
MLP use it. Everything else in the framework takes the cuBLAS path by default — this kernel has not been
tuned against it yet (one 128 x 256 tile per CTA, no persistence, no 2-CTA pairs).
"""
from __future__ import annotations

import os
from ctypes import c_int, c_void_p
from typing import Optional, Tuple

import torch

from ..runtime.native import NativeError, last_error, load_library

_bound = False
ACT = {"none": 0, "gelu": 1}


def _lib():
    global _bound
    lib = load_library()
    if not _bound:
        lib.adapcc_gemm_bias_act_v.argtypes = [c_void_p] * 5 + [c_int] * 5 + [c_void_p]
        _bound = True
    return lib


def supported(x: torch.Tensor, weight: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16
            and weight.shape[1] % 64 == 0 and weight.shape[0] % 128 == 0)


def default_variant() -> int:
    """0: one output tile per CTA (validated on B200). 1: persistent CTAs with a double-buffered TMEM accumulator
    (``ADAPCC_TCGEN05_VARIANT=1``; compiled and SASS-checked, first GPU run pending)."""
    return int(os.environ.get("ADAPCC_TCGEN05_VARIANT", "0"))


def linear_act(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], act: str = "gelu",
               save_pre: bool = False, variant: Optional[int] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """-> (act(x @ weight.T + bias), pre-activation or None). x [..., K], weight [N, K], bias [N]; bf16."""
    if not supported(x, weight):
        raise NativeError("linear_act: needs bf16 CUDA tensors with K % 64 == 0 and N % 128 == 0")
    n, k = weight.shape
    x2 = x.reshape(-1, k).contiguous()
    w = weight.contiguous()
    m = x2.shape[0]
    out = torch.empty(m, n, dtype=torch.bfloat16, device=x.device)
    pre = torch.empty_like(out) if save_pre else None
    b = bias.contiguous() if bias is not None else None
    rc = _lib().adapcc_gemm_bias_act_v(c_void_p(x2.data_ptr()), c_void_p(w.data_ptr()),
                                       c_void_p(b.data_ptr() if b is not None else 0), c_void_p(out.data_ptr()),
                                       c_void_p(pre.data_ptr() if pre is not None else 0), m, n, k, ACT[act],
                                       default_variant() if variant is None else int(variant),
                                       c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise NativeError(f"gemm_bias_act failed: {last_error()}")
    shape = x.shape[:-1] + (n,)
    return out.view(shape), (pre.view(shape) if pre is not None else None)


class _LinearGeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        out, pre = linear_act(x, w, b, "gelu", save_pre=True)
        ctx.save_for_backward(x, w, pre)
        ctx.params = (w, b)
        return out

    @staticmethod
    def backward(ctx, dy):
        from .layers import linear_backward
        x, w, pre = ctx.saved_tensors
        du = torch.ops.aten.gelu_backward(dy.contiguous(), pre, approximate="tanh")
        return linear_backward(ctx.params, x, w, du, ctx.needs_input_grad[0])


def linear_gelu(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """gelu_tanh(x @ weight.T + bias) with the bias and activation in the tcgen05 GEMM's epilogue."""
    return _LinearGeluFn.apply(x, weight, bias)
