"""GPT-2 small with the "double heads" (LM + multiple choice) of the reference's flagship workload.

The reference trains ``transformers.GPT2DoubleHeadsModel(GPT2Config())`` (12 layers, d=768, 12 heads,
ctx 1024, vocab 50257 + 5 special tokens) on PersonaChat-shaped batches
``input_ids/token_type_ids/lm_labels [B, C, T]``, ``mc_token_ids [B, C]``, ``mc_labels [B]`` with
loss = lm_coef * LM cross-entropy + mc_coef * multiple-choice cross-entropy
(/root/reference/models/gpt2/train_gpt2_ddp.py:28-31,100-107,157-195). This is the same
architecture written for B200 training:

* parameters live in bf16 (fp32 master copy inside the fused optimizer), so gradients are born in
  bf16 and the gradient all-reduce moves half the bytes with fp32 accumulation in the kernel;
* attention goes through ``scaled_dot_product_attention`` (flash kernels), MLP through cuBLAS;
* the LM head + cross-entropy is one fused pass per row chunk: bf16 logits from the GEMM, a softmax-CE
  kernel that overwrites them in place with their gradient, and the two backward GEMMs right away — fp32
  logits / probabilities (50 304 x 8192 x 4 B = 1.6 GB each in the stock path) never exist;
* no data-dependent host syncs: the whole step can be captured in one CUDA graph.

Token-type embeddings reuse ``wte`` exactly as HF GPT-2 does.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..ops.layers import FusedLayerNorm, FusedLinear


@dataclass
class GPT2Config:
    """GPT-2 small hyper-parameters (HF ``GPT2Config()`` + the five dialogue tokens) — the model of the reference's
    flagship workload, /root/reference/models/gpt2/train_gpt2_ddp.py:157-159."""

    vocab_size: int = 50257 + 5          # GPT2Config() + the five PersonaChat special tokens
    n_positions: int = 1024
    n_embd: int = 768
    n_layer: int = 12
    n_head: int = 12
    layer_norm_epsilon: float = 1e-5
    initializer_range: float = 0.02
    lm_chunk_rows: int = 8192            # rows of the fused LM-head/CE evaluated at a time (bf16 logits of one
                                         # chunk: 8192 x 50304 x 2 B = 0.8 GB; measured 2 % faster per step than 2048)

    @classmethod
    def tiny(cls) -> "GPT2Config":
        return cls(vocab_size=512, n_positions=64, n_embd=64, n_layer=2, n_head=4, lm_chunk_rows=64)


class Block(nn.Module):
    """Pre-LN transformer block; the MLP runs as two GEMMs with the activation passes in tcgen05 epilogues when
    supported (ops/gemm.py), residual adds optionally fused into the following LayerNorm (``forward_deferred``)."""

    def __init__(self, cfg: GPT2Config):
        super().__init__()
        d = cfg.n_embd
        self.n_head = cfg.n_head
        # fused kernels on CUDA/bf16, stock ATen otherwise (same parameters either way)
        self.ln_1 = FusedLayerNorm(d, eps=cfg.layer_norm_epsilon)
        self.c_attn = FusedLinear(d, 3 * d)
        self.c_proj = FusedLinear(d, d)
        self.ln_2 = FusedLayerNorm(d, eps=cfg.layer_norm_epsilon)
        self.c_fc = FusedLinear(d, 4 * d)
        self.c_proj2 = FusedLinear(4 * d, d)
        # the MLP's activation passes live in our tcgen05 GEMM's epilogues (ops/gemm.py, csrc/gemm_tcgen05_pp.cu):
        # 2 (default): c_fc + bias + GELU forward AND dGELU + the c_fc bias gradient in the backward GEMM dY.W2 — no
        #    stand-alone GELU / GELU-backward / column-sum kernel (measured -0.32 ms per GPT-2 step on B200);
        # 1: forward fusion only; 0: cuBLAS + separate activation kernels
        self.tc_mlp = int(os.environ.get("ADAPCC_TCGEN05_MLP", "2") or 0)

    def _mlp(self, h: torch.Tensor) -> torch.Tensor:
        if self.tc_mlp and h.is_cuda and h.dtype == torch.bfloat16 and torch.is_grad_enabled():
            from ..ops.gemm import linear_gelu, mlp_gelu, supported
            if supported(h, self.c_fc.weight):
                if self.tc_mlp >= 2 and self.c_fc.weight.shape[0] % 64 == 0:
                    return mlp_gelu(h, self.c_fc.weight, self.c_fc.bias, self.c_proj2.weight, self.c_proj2.bias)
                return self.c_proj2(linear_gelu(h, self.c_fc.weight, self.c_fc.bias))
        return self.c_proj2(F.gelu(self.c_fc(h), approximate="tanh"))

    def _attn(self, h: torch.Tensor) -> torch.Tensor:
        B, T, D = h.shape
        # split along the feature dim (views; measured 6 % faster fwd+bwd than the packed permute:
        # the backward writes dq/dk/dv straight into one [B, T, 3D] buffer layout-wise)
        q, k, v = (t.view(B, T, self.n_head, D // self.n_head).transpose(1, 2) for t in self.c_attn(h).split(D, dim=-1))
        a = F.scaled_dot_product_attention(q, k, v, is_causal=True)
        return self.c_proj(a.transpose(1, 2).reshape(B, T, D))

    def forward_deferred(self, x: torch.Tensor, delta: Optional[torch.Tensor]):
        """Same block with every residual add fused into the LayerNorm that follows it: takes the
        residual stream ``x`` and the previous block's not-yet-added MLP output ``delta``, returns
        (stream, this block's not-yet-added MLP output)."""
        if delta is None:
            h = self.ln_1(x)
        else:
            x, h = self.ln_1.forward_add(x, delta)
        x, h = self.ln_2.forward_add(x, self._attn(h))
        return x, self._mlp(h)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x + self._attn(self.ln_1(x))
        return x + self._mlp(self.ln_2(x))


def use_sink(p) -> bool:
    from ..ops.layers import _sink
    return p.is_cuda and _sink(p) is not None


class _ChunkedLMLoss(torch.autograd.Function):
    """sum over rows of CE(h @ W^T, labels), ``chunk`` rows at a time. The backward of every chunk
    is produced inside the forward pass (fused softmax-CE kernel overwrites the bf16 logits with
    their gradient), so logits never outlive a chunk and are touched twice instead of ~8 times."""

    @staticmethod
    def forward(ctx, h, weight, labels, chunk, vocab, scale=None):
        """``scale`` (1-element fp32 device tensor or None): the result is ``scale * sum of row losses`` and the
        gradients produced here already contain it — the caller promises that the returned value enters the total
        loss with coefficient exactly 1 (GPT2DoubleHeads.forward does). That is what allows the weight gradient to be
        written straight into the engine's flat gradient buffer during the forward pass (``weight`` carries a grad sink,
        ops/layers.py::_sink): no 77 MB ``grad_w * g`` and no 77 MB accumulate kernel in the backward."""
        n = h.shape[0]
        use_kernel = h.is_cuda and h.dtype == torch.bfloat16 and weight.shape[0] % 8 == 0
        grad_h = torch.empty_like(h)
        grad_w = None
        total = torch.zeros((), dtype=torch.float32, device=h.device)
        direct = False
        if use_kernel:
            from ..ops import fused_ce_
            from ..ops.layers import _sink
            direct = (scale is not None and _sink(weight) is not None and weight.grad.dtype == h.dtype
                      and ctx.needs_input_grad[1])
        for s in range(0, n, chunk):
            hs = h[s:s + chunk]
            ls = labels[s:s + chunk]
            logits = hs @ weight.t()
            if use_kernel:
                total += fused_ce_(logits, ls, vocab, grad_scale=scale).sum()
                g = logits                                         # now d loss / d logits (bf16)
            else:                                                  # CPU / fp32 reference path
                lf = logits.float()
                lf[:, vocab:] = float("-inf")
                valid = (ls >= 0)
                lse = torch.logsumexp(lf, dim=-1)
                tgt = lf.gather(1, ls.clamp(min=0).unsqueeze(1)).squeeze(1)
                total += ((lse - tgt) * valid).sum()
                p = torch.softmax(lf, dim=-1)
                p.scatter_add_(1, ls.clamp(min=0).unsqueeze(1), -torch.ones_like(p[:, :1]))
                g = p * valid.unsqueeze(1)
                if scale is not None:
                    g = g * scale
                g = g.to(h.dtype)
            torch.mm(g, weight, out=grad_h[s:s + chunk])
            # dW accumulates across chunks inside the GEMM (fp32 accumulator, beta = 1 epilogue) — an
            # fp32 side buffer cost ~1 GB of extra HBM traffic per chunk at vocab 50304
            if direct:
                if s == 0:
                    torch.mm(g.t(), hs, out=weight.grad)           # the flat-buffer view (zeroed at step start)
                else:
                    weight.grad.addmm_(g.t(), hs)
            elif grad_w is None:
                grad_w = g.t() @ hs
            else:
                grad_w.addmm_(g.t(), hs)
        if scale is not None:
            total = total * scale.reshape(())
        ctx.direct, ctx.weight = direct, weight
        if direct:
            ctx.save_for_backward(grad_h)
        else:
            if grad_w is None:
                grad_w = torch.zeros_like(weight)
            ctx.save_for_backward(grad_h, grad_w)
        return total

    @staticmethod
    def backward(ctx, g):
        if ctx.direct:                       # unit upstream gradient by contract; dW is already in weight.grad
            return ctx.saved_tensors[0], None, None, None, None, None
        grad_h, grad_w = ctx.saved_tensors
        gw = grad_w * g.to(grad_w.dtype)
        if use_sink(ctx.weight):
            # the table's gradient is owned by a sink (its LAST writer, the fused embedding backward, reports it):
            # autograd must not see a gradient for it, or the engine would count the parameter twice
            ctx.weight.grad.add_(gw)
            gw = None
        return grad_h * g.to(grad_h.dtype), gw, None, None, None, None


class GPT2DoubleHeads(nn.Module):
    """GPT-2 with a tied LM head and a multiple-choice head on a chosen token (HF ``GPT2DoubleHeadsModel``):
    ``forward`` returns the training losses (LM on labelled positions + MC over candidates); fused embedding,
    chunked LM-head + CE, optional scored-rows-only LM head."""

    def __init__(self, cfg: Optional[GPT2Config] = None):
        super().__init__()
        self.cfg = cfg = cfg or GPT2Config()
        # rows padded to a multiple of 64 so the LM-head GEMM and the fused CE kernel see aligned rows;
        # ids >= vocab_size are never produced and their logits are masked out
        self.padded_vocab = (cfg.vocab_size + 63) // 64 * 64
        self.wte = nn.Embedding(self.padded_vocab, cfg.n_embd)
        self.wpe = nn.Embedding(cfg.n_positions, cfg.n_embd)
        self.h = nn.ModuleList([Block(cfg) for _ in range(cfg.n_layer)])
        self.ln_f = FusedLayerNorm(cfg.n_embd, eps=cfg.layer_norm_epsilon)
        self.mc_head = nn.Linear(cfg.n_embd, 1)           # SequenceSummary(summary_type="cls_index")
        # > 0: evaluate the LM head only on (at most) this many rows whose label is not -100. Rows with
        # an ignored label contribute neither loss nor gradient, so the result is identical; PersonaChat
        # batches score only the last candidate's reply, i.e. ~1/8 of the rows at the bench shape.
        self.lm_row_capacity = 0
        # residual adds fused into the following LayerNorm (fwd) / its input gradient (bwd)
        # (default on since the fused tcgen05 MLP: A/B in one gpurun call 8.53 vs 8.58-8.66 ms per step)
        self.fuse_add_ln = os.environ.get("ADAPCC_FUSE_ADD_LN", "1") == "1"
        # wte[ids] + wpe[pos] + wte[token types] as one kernel, sort-free fp32-accumulating backward
        # (csrc/ops_embed.cu); the flat engine gives both tables a gradient sink (parallel/engine.py)
        self.fused_embed = os.environ.get("ADAPCC_FUSED_EMBED", "1") != "0"
        if self.fused_embed:
            self.wte.weight._adapcc_embed_table = True
            self.wpe.weight._adapcc_embed_table = True
        self._pos_cache = {}
        self.apply(self._init)
        for blk in self.h:                                  # GPT-2 residual-projection scaling
            for lin in (blk.c_proj, blk.c_proj2):
                nn.init.normal_(lin.weight, std=cfg.initializer_range / math.sqrt(2 * cfg.n_layer))

    def _init(self, m):
        if isinstance(m, (nn.Linear, nn.Embedding)):
            nn.init.normal_(m.weight, std=self.cfg.initializer_range)
            if isinstance(m, nn.Linear) and m.bias is not None:
                nn.init.zeros_(m.bias)

    def num_parameters(self) -> int:
        return sum(p.numel() for p in self.parameters())

    def hidden(self, input_ids: torch.Tensor, token_type_ids: Optional[torch.Tensor]) -> torch.Tensor:
        N, T = input_ids.shape
        if (self.fused_embed and input_ids.is_cuda and self.wte.weight.dtype == torch.bfloat16
                and self.wpe.weight.dtype == torch.bfloat16):
            from ..ops import fused_embedding_sum
            key = (N, T, input_ids.device)
            pos = self._pos_cache.get(key)
            if pos is None:
                pos = self._pos_cache[key] = torch.arange(T, device=input_ids.device).repeat(N)
            lookups = [(0, input_ids), (1, pos)] + ([(0, token_type_ids)] if token_type_ids is not None else [])
            x = fused_embedding_sum([self.wte.weight, self.wpe.weight], lookups).view(N, T, -1)
        else:
            pos = torch.arange(T, device=input_ids.device)
            x = self.wte(input_ids) + self.wpe(pos)[None]
            if token_type_ids is not None:
                x = x + self.wte(token_type_ids)
        if self.fuse_add_ln:
            delta = None
            for blk in self.h:
                x, delta = blk.forward_deferred(x, delta)
            return self.ln_f.forward_add(x, delta)[1]
        for blk in self.h:
            x = blk(x)
        return self.ln_f(x)

    def forward(self, input_ids, token_type_ids=None, mc_token_ids=None, lm_labels=None, mc_labels=None,
                lm_coef: float = 1.0, mc_coef: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """-> (loss, lm_loss, mc_loss). Shapes as in the reference: [B, C, T] ids/labels, [B, C]
        mc_token_ids, [B] mc_labels; -100 labels are ignored."""
        B, C, T = input_ids.shape
        ids = input_ids.reshape(B * C, T)
        tt = token_type_ids.reshape(B * C, T) if token_type_ids is not None else None
        h = self.hidden(ids, tt)                                              # [B*C, T, D]
        zero = torch.zeros((), dtype=torch.float32, device=h.device)
        lm_loss, mc_loss = zero, zero
        if lm_labels is not None:
            labels = lm_labels.reshape(B * C, T)
            shift_h = h[:, :-1].reshape(-1, h.shape[-1])
            shift_l = labels[:, 1:].reshape(-1)
            n_valid = (shift_l >= 0).sum()
            cap = int(self.lm_row_capacity)
            compact = 0 < cap < shift_l.numel()
            if compact:
                # scored rows first, in their original order (stable sort; static shapes, so the step
                # stays CUDA-graph capturable). Backward scatters the row gradients back.
                order = torch.argsort((shift_l < 0).to(torch.int8), stable=True)[:cap]
                shift_h = shift_h.index_select(0, order)
                shift_l = shift_l.index_select(0, order)
            # the mean's 1/n and lm_coef travel INTO the fused CE kernel, so the LM term enters `loss` with
            # coefficient 1 (the contract of _ChunkedLMLoss's direct weight-gradient path)
            scale = (lm_coef / n_valid.clamp(min=1).to(torch.float32)).reshape(1)
            lm_term = _ChunkedLMLoss.apply(shift_h, self.wte.weight, shift_l, self.cfg.lm_chunk_rows,
                                           self.cfg.vocab_size, scale)
            lm_loss = lm_term / lm_coef if lm_coef != 0 else lm_term
            if compact:
                # a batch with more scored rows than the capacity must not train silently on a subset
                lm_loss = torch.where(n_valid > cap, torch.full_like(lm_loss, float("nan")), lm_loss)
                lm_term = torch.where(n_valid > cap, torch.full_like(lm_term, float("nan")), lm_term)
        if mc_token_ids is not None and mc_labels is not None:
            idx = mc_token_ids.reshape(B * C, 1, 1).expand(-1, 1, h.shape[-1])
            cls_h = h.gather(1, idx).squeeze(1)                               # [B*C, D]
            mc_logits = self.mc_head(cls_h).view(B, C).float()
            mc_loss = F.cross_entropy(mc_logits, mc_labels)
        loss = (lm_term if lm_labels is not None else zero) + mc_coef * mc_loss
        return loss, lm_loss, mc_loss


@torch.no_grad()
def sample_sequence(model: "GPT2DoubleHeads", input_ids: torch.Tensor, token_type_ids: Optional[torch.Tensor] = None,
                    max_new_tokens: int = 20, temperature: float = 0.7, top_k: int = 0, top_p: float = 0.9,
                    eos_token: Optional[int] = None, reply_type: Optional[int] = None,
                    generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Nucleus / top-k sampling of a reply, one token at a time (the decoding loop of the reference's interact.py:
    ``top_filtering`` + ``sample_sequence``, /root/reference/models/gpt2/interact.py). ``input_ids`` [1, T]; new tokens
    get ``reply_type`` as their token type when given. No KV cache: a demo / evaluation path, not a serving engine."""
    ids = input_ids.clone()
    tt = token_type_ids.clone() if token_type_ids is not None else None
    vocab = model.cfg.vocab_size
    for _ in range(max_new_tokens):
        window = slice(max(0, ids.shape[1] - model.cfg.n_positions), None)
        h = model.hidden(ids[:, window], tt[:, window] if tt is not None else None)
        logits = (h[:, -1].float() @ model.wte.weight.float().t())[:, :vocab] / max(temperature, 1e-6)
        if top_k > 0:
            kth = torch.topk(logits, min(top_k, vocab)).values[:, -1:]
            logits = logits.masked_fill(logits < kth, float("-inf"))
        if 0.0 < top_p < 1.0:
            srt, idx = torch.sort(logits, descending=True)
            cum = torch.softmax(srt, -1).cumsum(-1)
            drop = cum > top_p
            drop[:, 1:] = drop[:, :-1].clone()                  # keep the first token that crosses the threshold
            drop[:, 0] = False
            logits = logits.masked_fill(torch.zeros_like(drop).scatter(1, idx, drop), float("-inf"))
        nxt = torch.multinomial(torch.softmax(logits, -1), 1, generator=generator)
        ids = torch.cat([ids, nxt], 1)
        if tt is not None:
            tt = torch.cat([tt, torch.full_like(nxt, reply_type if reply_type is not None else int(tt[0, -1]))], 1)
        if eos_token is not None and int(nxt) == eos_token:
            break
    return ids[:, input_ids.shape[1]:]


def lm_rows_needed(lm_labels: torch.Tensor, multiple: int = 256) -> int:
    """Capacity for ``GPT2DoubleHeads.lm_row_capacity`` from a (host) label tensor [B, C, T]: the number
    of scored next-token rows, rounded up to ``multiple``."""
    n = int((lm_labels[..., 1:] >= 0).sum())
    return max(multiple, (n + multiple - 1) // multiple * multiple)


def synthetic_batch(batch: int, candidates: int, seq_len: int, vocab: int, device="cpu", seed: int = 0,
                    pin: bool = False):
    """PersonaChat-shaped synthetic batch (the real set is downloaded by the reference; there is no
    network here): history tokens carry -100 labels, only the last candidate's reply is scored."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, vocab, (batch, candidates, seq_len), generator=g)
    tt = torch.randint(vocab - 5, vocab - 3, (batch, candidates, seq_len), generator=g)
    labels = torch.full((batch, candidates, seq_len), -100, dtype=torch.long)
    reply = max(1, seq_len // 4)
    labels[:, -1, -reply:] = ids[:, -1, -reply:]
    mc_token_ids = torch.full((batch, candidates), seq_len - 1, dtype=torch.long)
    mc_labels = torch.full((batch,), candidates - 1, dtype=torch.long)
    out = {"input_ids": ids, "token_type_ids": tt, "lm_labels": labels, "mc_token_ids": mc_token_ids,
           "mc_labels": mc_labels}
    if pin and torch.cuda.is_available():
        out = {k: v.pin_memory() for k, v in out.items()}
    if device != "cpu":
        out = {k: v.to(device, non_blocking=True) for k, v in out.items()}
    return out
