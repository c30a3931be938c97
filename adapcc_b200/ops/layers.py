"""nn.Module wrappers around the fused transformer kernels (csrc/ops_norm.cu).

* ``FusedLayerNorm`` — single-pass bf16 LayerNorm forward/backward (row kept in registers, dgamma /
  dbeta reduced in the same pass). Falls back to ``F.layer_norm`` for CPU tensors, non-bf16 dtypes
  or widths other than 256/512/768/1024.
* ``FusedLinear`` — ``F.linear`` (cuBLASLt, bias epilogue) whose backward computes the bias gradient
  with the column-sum kernel instead of ATen's generic reduction.
"""
from __future__ import annotations

from ctypes import c_float, c_int, c_void_p

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..runtime.native import NativeError, last_error, load_library

_bound = False
_LN_WIDTHS = (256, 512, 768, 1024)


def _lib():
    global _bound
    lib = load_library()
    if not _bound:
        lib.adapcc_ln_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                      c_float, c_void_p]
        lib.adapcc_ln_bwd.argtypes = [c_void_p] * 9 + [c_int, c_int, c_void_p]
        lib.adapcc_ln_partials.argtypes = [c_int]
        lib.adapcc_colsum.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]
        lib.adapcc_colsum_splits.argtypes = [c_int]
        _bound = True
    return lib


def _s():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return c_void_p(t.data_ptr())


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        lib = _lib()
        d = x.shape[-1]
        x2 = x.reshape(-1, d).contiguous()
        rows = x2.shape[0]
        y = torch.empty_like(x2)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        if lib.adapcc_ln_fwd(_p(x2), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd), rows, d, eps, _s()) != 0:
            raise NativeError(f"ln_fwd failed: {last_error()}")
        ctx.save_for_backward(x2, gamma, mean, rstd)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        lib = _lib()
        x2, gamma, mean, rstd = ctx.saved_tensors
        rows, d = x2.shape
        dy2 = dy.reshape(rows, d).contiguous()
        dx = torch.empty_like(x2)
        dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(gamma)
        part = torch.empty(2 * lib.adapcc_ln_partials(rows) * d, dtype=torch.float32, device=x2.device)
        if lib.adapcc_ln_bwd(_p(dy2), _p(x2), _p(gamma), _p(mean), _p(rstd), _p(dx), _p(dgamma), _p(dbeta), _p(part),
                             rows, d, _s()) != 0:
            raise NativeError(f"ln_bwd failed: {last_error()}")
        return dx.view(dy.shape), dgamma, dbeta, None


class FusedLayerNorm(nn.LayerNorm):
    def forward(self, x):
        if (x.is_cuda and x.dtype == torch.bfloat16 and self.weight.dtype == torch.bfloat16
                and x.shape[-1] in _LN_WIDTHS and len(self.normalized_shape) == 1):
            return _LayerNormFn.apply(x, self.weight, self.bias, self.eps)
        return F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return F.linear(x, w, b)

    @staticmethod
    def backward(ctx, dy):
        lib = _lib()
        x, w = ctx.saved_tensors
        n = w.shape[0]
        dy2 = dy.reshape(-1, n)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        x2 = x.reshape(-1, x.shape[-1])
        dx = (dy2 @ w).view(x.shape) if ctx.needs_input_grad[0] else None
        dw = dy2.t() @ x2
        rows = dy2.shape[0]
        db = torch.empty(n, dtype=dy2.dtype, device=dy2.device)
        part = torch.empty(lib.adapcc_colsum_splits(rows) * n, dtype=torch.float32, device=dy2.device)
        if lib.adapcc_colsum(_p(dy2), rows, n, _p(db), _p(part), _s()) != 0:
            raise NativeError(f"colsum failed: {last_error()}")
        return dx, dw, db


class FusedLinear(nn.Linear):
    def forward(self, x):
        if (x.is_cuda and x.dtype == torch.bfloat16 and self.bias is not None and self.weight.dtype == torch.bfloat16
                and self.out_features % 8 == 0 and torch.is_grad_enabled()):
            return _LinearFn.apply(x, self.weight, self.bias)
        return F.linear(x, self.weight, self.bias)
