"""Hand-written sm_100a ops used by the training engine and the workloads (bound through the same
in-tree ``libadapcc.so`` as the collectives). Every function launches on the current CUDA stream
and fails loudly if the native library is missing — there is no silent PyTorch fallback on a GPU.
"""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int, c_longlong, c_void_p
from typing import Optional

import torch

from ..constants import DTYPE_IDS
from ..runtime.native import NativeError, last_error, load_library

_bound = False


def _lib():
    global _bound
    lib = load_library()
    if not _bound:
        lib.adapcc_sumsq.argtypes = [c_void_p, c_longlong, c_int, c_void_p, c_void_p]
        lib.adapcc_fused_adamw.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_int,
                                           c_float, c_float, c_float, c_float, c_float, c_int, c_float, c_float,
                                           c_void_p, c_void_p, c_void_p]
        lib.adapcc_fused_adamw_lr.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_int,
                                              c_float, c_float, c_float, c_float, c_float, c_int, c_float, c_float,
                                              c_void_p, c_void_p, c_void_p, c_void_p]
        lib.adapcc_fused_sgd.argtypes = [c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_int, c_float, c_float,
                                         c_void_p]
        lib.adapcc_incr_int.argtypes = [c_void_p, c_void_p]
        lib.adapcc_zero_adamw_bcast.argtypes = [c_void_p, c_void_p, c_int, c_longlong, c_void_p, c_void_p, c_void_p,
                                                c_void_p, c_longlong] + [c_float] * 7 + [c_void_p, c_void_p, c_void_p]
        lib.adapcc_fused_ce.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]
        lib.adapcc_fused_ce_scaled.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]
        lib.adapcc_embed_sum_fwd.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]
        lib.adapcc_embed_sum_bwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                             c_void_p, c_void_p]
        lib.adapcc_embed_bwd_reset.argtypes = [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p]
        _bound = True
    return lib


def _stream() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _dt(t: torch.Tensor) -> int:
    return DTYPE_IDS[str(t.dtype).replace("torch.", "")]


def _ck(rc: int, what: str) -> None:
    if rc != 0:
        raise NativeError(f"{what} failed: {last_error()}")


def sumsq_(grad: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out (1-element fp32, already zeroed) += sum(grad**2)."""
    _ck(_lib().adapcc_sumsq(c_void_p(grad.data_ptr()), grad.numel(), _dt(grad), c_void_p(out.data_ptr()), _stream()),
        "sumsq")
    return out


def fused_adamw_(param: torch.Tensor, grad: torch.Tensor, master: torch.Tensor, m: torch.Tensor, v: torch.Tensor, *,
                 lr: float, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01, step: int = 1,
                 max_norm: float = 0.0, grad_scale: float = 1.0, sumsq: Optional[torch.Tensor] = None,
                 step_tensor: Optional[torch.Tensor] = None, lr_tensor: Optional[torch.Tensor] = None) -> None:
    """One launch: clip (coefficient derived on the device from ``sumsq``), AdamW on the fp32
    master/m/v, write-back of the (bf16 or fp32) parameters. ``step_tensor`` (int32 on device) makes
    the bias correction replayable inside a CUDA graph; ``lr_tensor`` (fp32 on device) does the same for the learning
    rate, so a schedule only has to update that scalar between replays."""
    _ck(_lib().adapcc_fused_adamw_lr(c_void_p(param.data_ptr()), c_void_p(grad.data_ptr()), c_void_p(master.data_ptr()),
                                     c_void_p(m.data_ptr()), c_void_p(v.data_ptr()), param.numel(), _dt(param), _dt(grad),
                                     lr, betas[0], betas[1], eps, weight_decay, int(step), max_norm, grad_scale,
                                     c_void_p(sumsq.data_ptr() if sumsq is not None else None),
                                     c_void_p(step_tensor.data_ptr() if step_tensor is not None else None),
                                     c_void_p(lr_tensor.data_ptr() if lr_tensor is not None else None), _stream()),
        "fused_adamw")


def zero_adamw_bcast_(comm, param_shard: torch.Tensor, grad_shard: torch.Tensor, master: torch.Tensor,
                      m: torch.Tensor, v: torch.Tensor, *, lr: float, betas=(0.9, 0.999), eps: float = 1e-8,
                      weight_decay: float = 0.01, max_norm: float = 0.0, grad_scale: float = 1.0,
                      sumsq: Optional[torch.Tensor] = None, step_tensor: torch.Tensor = None) -> None:
    """ZeRO-1 step for one shard (csrc/zero.cu): AdamW on this rank's slice of the (symmetric-heap) parameter
    buffer; the updated bf16 parameters are written through the heap's multicast alias (``multimem.st``: the switch
    replicates them into every rank's copy) or, without multicast, stored into every peer's buffer. ``param_shard``
    must be a view inside ``comm``'s symmetric heap; master / m / v are the shard's fp32 state. The engine's default at
    N > 1 (``tests/gpu_zero1_worker.py``: parity with the replicated optimizer at 2 and 8 GPUs)."""
    n = param_shard.numel()
    if n == 0:
        return
    off = comm.heap_offset(param_shard)
    mc = comm.heap_mc_ptr()
    peers = None
    npeers = 0
    if not mc:
        ptrs = [comm.peer_heap_ptr(r) for r in range(comm.world)]
        peers = (c_void_p * len(ptrs))(*ptrs)
        npeers = len(ptrs)
    _ck(_lib().adapcc_zero_adamw_bcast(c_void_p(mc or None), peers, npeers, off, c_void_p(grad_shard.data_ptr()),
                                       c_void_p(master.data_ptr()), c_void_p(m.data_ptr()), c_void_p(v.data_ptr()), n,
                                       lr, betas[0], betas[1], eps, weight_decay, max_norm, grad_scale,
                                       c_void_p(sumsq.data_ptr() if sumsq is not None else None),
                                       c_void_p(step_tensor.data_ptr()), _stream()), "zero_adamw_bcast")


def fused_sgd_(param, grad, master, *, lr: float, grad_scale: float = 1.0) -> None:
    _ck(_lib().adapcc_fused_sgd(c_void_p(param.data_ptr()), c_void_p(grad.data_ptr()), c_void_p(master.data_ptr()),
                                param.numel(), _dt(param), _dt(grad), lr, grad_scale, _stream()), "fused_sgd")


def incr_(t: torch.Tensor) -> None:
    _ck(_lib().adapcc_incr_int(c_void_p(t.data_ptr()), _stream()), "incr")


def fused_ce_(logits: torch.Tensor, labels: torch.Tensor, vocab: int,
              grad_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """In place: bf16 ``logits`` [rows, stride] become d(sum of row losses)/d logits; returns the
    per-row losses (fp32). Rows whose label is negative are ignored (loss 0, zero gradient).
    ``grad_scale`` (1-element fp32 on the device): the gradient is multiplied by it inside the kernel (the row losses
    are not), so a mean loss's 1/n needs no extra pass over the gradient or over the GEMMs fed by it."""
    assert logits.dtype == torch.bfloat16 and logits.is_contiguous() and logits.dim() == 2
    rows, stride = logits.shape
    row_loss = torch.empty(rows, dtype=torch.float32, device=logits.device)
    gs = None
    if grad_scale is not None:
        assert grad_scale.dtype == torch.float32 and grad_scale.numel() == 1 and grad_scale.is_cuda
        gs = grad_scale.data_ptr()
    _ck(_lib().adapcc_fused_ce_scaled(c_void_p(logits.data_ptr()), c_void_p(labels.data_ptr()),
                                      c_void_p(row_loss.data_ptr()), rows, int(vocab), stride, c_void_p(gs), _stream()),
        "fused_ce")
    return row_loss


# ---- fused embedding sum (csrc/ops_embed.cu) ----------------------------------------------------------------------
_embed_work = {}


def _embed_buffers(device, n_owner: int, scratch_elems: int):
    """(owner int32 [n_owner] all INT_MAX, scratch fp32 [scratch_elems] all zero): the backward kernels leave both in
    exactly that state again, so they are initialised once per (device, size)."""
    key = (device.index, n_owner, scratch_elems)
    buf = _embed_work.get(key)
    if buf is None:
        owner = torch.empty(n_owner, dtype=torch.int32, device=device)
        scratch = torch.empty(scratch_elems, dtype=torch.float32, device=device)
        _ck(_lib().adapcc_embed_bwd_reset(c_void_p(owner.data_ptr()), n_owner, c_void_p(scratch.data_ptr()),
                                          scratch_elems, _stream()), "embed_bwd_reset")
        buf = _embed_work[key] = (owner, scratch)
    return buf


def _ptr_array(tensors):
    return (c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


class _EmbedSumFn(torch.autograd.Function):
    """y[t] = sum_k tables[spec[k]][idx_k[t]] (bf16, fp32 sum); args: spec, K index tensors, then the tables."""

    @staticmethod
    def forward(ctx, spec, *args):
        K = len(spec)
        idx, tables = args[:K], args[K:]
        n, D = idx[0].numel(), tables[0].shape[1]
        out = torch.empty(n, D, dtype=tables[0].dtype, device=tables[0].device)
        _ck(_lib().adapcc_embed_sum_fwd(_ptr_array([tables[j] for j in spec]), _ptr_array(idx), K, n, D,
                                        c_void_p(out.data_ptr()), _stream()), "embed_sum_fwd")
        ctx.spec, ctx.idx, ctx.tables, ctx.n, ctx.D = spec, idx, tables, n, D
        return out

    @staticmethod
    def backward(ctx, dy):
        from .layers import _sink

        spec, idx, tables, n, D = ctx.spec, ctx.idx, ctx.tables, ctx.n, ctx.D
        K = len(spec)
        dy = dy.contiguous()
        base, off = [], 0
        for t in tables:                                   # owner slots: one range per distinct table
            base.append(off)
            off += t.shape[0]
        owner, scratch = _embed_buffers(dy.device, off, K * n * D)
        sinks = [_sink(t) for t in tables]
        grads, ret = [], []
        for t, s in zip(tables, sinks):
            if s is not None and s.begin():                # straight into the engine's flat gradient view (which may
                grads.append(t.grad)                       # already hold another contribution, e.g. a tied LM head)
                ret.append(None)
            else:
                g = torch.zeros_like(t)
                grads.append(g)
                ret.append(g)
        ob = (c_int * K)(*[base[j] for j in spec])
        _ck(_lib().adapcc_embed_sum_bwd(c_void_p(dy.data_ptr()), _ptr_array(idx), _ptr_array([grads[j] for j in spec]),
                                        ob, K, n, D, c_void_p(owner.data_ptr()), c_void_p(scratch.data_ptr()),
                                        _stream()), "embed_sum_bwd")
        for s in sinks:
            if s is not None:
                s.done()
        return (None,) + (None,) * K + tuple(ret)


def fused_embedding_sum(tables, lookups) -> torch.Tensor:
    """``sum_k tables[j_k][idx_k]`` for ``lookups = [(j_k, idx_k), ...]`` (idx_k: int64, all the same number of
    elements n) -> [n, D]. bf16 CUDA tables: one forward kernel and a sort-free three-launch backward that accumulates
    duplicate rows in fp32 (csrc/ops_embed.cu). Anything else: plain ``F.embedding`` sums."""
    t0 = tables[0]
    fused = (t0.is_cuda and all(t.dtype == torch.bfloat16 and t.is_contiguous() and t.shape[1] == t0.shape[1]
                                for t in tables) and t0.shape[1] % 8 == 0 and 1 <= len(lookups) <= 4)
    if not fused:
        out = None
        for j, ix in lookups:
            e = torch.nn.functional.embedding(ix.reshape(-1), tables[j])
            out = e if out is None else out + e
        return out
    spec = tuple(int(j) for j, _ in lookups)
    idx = tuple(ix.reshape(-1).contiguous() for _, ix in lookups)
    return _EmbedSumFn.apply(spec, *idx, *tables)
