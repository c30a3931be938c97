"""ConvAI2-style evaluation of the double-heads GPT-2: hits@1, perplexity, F1 — and the reply sampler they share.

The reference evaluates through ParlAI (``eval_hits`` / ``eval_ppl`` / ``eval_f1`` driving a ``TransformerAgent``,
/root/reference/models/gpt2/convai_evaluation.py:27-239), which is not available offline; the three metrics are
computed here directly on PersonaChat-schema data:

* **hits@1** — the multiple-choice head ranks the gold reply first among the utterance's candidates
  (convai_evaluation.py:122-146);
* **perplexity** — token-level ``exp(mean NLL)`` of the gold reply, and ParlAI's *word-level* variant built on
  :func:`next_word_probability` (a BPE-prefix → words table turns the model's next-token distribution into a
  distribution over dictionary words, convai_evaluation.py:160-197);
* **F1** — word-overlap F1 between a sampled reply and the gold one (ParlAI's normalisation: lower case, no
  punctuation, no articles).

Batched over candidates, no ParlAI, runs on CPU (tests) or GPU.
"""
from __future__ import annotations

import math
import re
import string
from collections import Counter, defaultdict
from typing import Dict, Iterable, List, Optional, Sequence

import torch
import torch.nn.functional as F

from ..data.personachat import IGNORE, build_input_from_segments
from ..data.tokenizer import DialogTokenizer


# ----------------------------------------------------------------------------------------------------------------
# model access (the training forward only returns losses)
# ----------------------------------------------------------------------------------------------------------------
def _lm_logits(model, h: torch.Tensor) -> torch.Tensor:
    return (h.float() @ model.wte.weight.float().t())[..., : model.cfg.vocab_size]


@torch.no_grad()
def mc_scores(model, input_ids: torch.Tensor, token_type_ids: torch.Tensor, mc_token_ids: torch.Tensor) -> torch.Tensor:
    """Multiple-choice logits ``[B, C]`` for ``[B, C, T]`` inputs."""
    B, C, T = input_ids.shape
    h = model.hidden(input_ids.reshape(B * C, T), token_type_ids.reshape(B * C, T))
    idx = mc_token_ids.reshape(B * C, 1, 1).expand(-1, 1, h.shape[-1])
    return model.mc_head(h.gather(1, idx).squeeze(1)).view(B, C).float()


# ----------------------------------------------------------------------------------------------------------------
# sampling
# ----------------------------------------------------------------------------------------------------------------
def top_filtering(logits: torch.Tensor, top_k: int = 0, top_p: float = 0.9, threshold: float = -float("inf")) -> torch.Tensor:
    """Top-k, nucleus and absolute-threshold filtering of a 1-D logit vector (returns a new tensor)."""
    logits = logits.clone()
    if top_k > 0:
        kth = torch.topk(logits, min(top_k, logits.numel())).values[-1]
        logits[logits < kth] = -float("inf")
    if top_p > 0.0:
        srt, idx = torch.sort(logits, descending=True)
        beyond = torch.cumsum(F.softmax(srt, dim=-1), dim=-1) > top_p
        beyond = torch.cat([beyond.new_zeros(1), beyond[:-1]])          # keep the token that crosses the threshold
        logits[idx[beyond]] = -float("inf")
    logits[logits < threshold] = -float("inf")
    return logits


@torch.no_grad()
def sample_reply(model, tokenizer: DialogTokenizer, persona, history, max_length: int = 20, min_length: int = 1,
                 temperature: float = 0.7, top_k: int = 0, top_p: float = 0.9, no_sample: bool = False,
                 device="cpu", generator: Optional[torch.Generator] = None) -> List[int]:
    """Generate a reply token by token: the input is rebuilt from persona + history + the partial reply each step, so
    speaker tokens and token types stay those of training (the reference's ``sample_sequence``, interact.py:59-87);
    special tokens end the reply, and are re-drawn while the reply is shorter than ``min_length``."""
    special = set(tokenizer.special_ids)
    out: List[int] = []
    for step in range(max_length):
        inst = build_input_from_segments(persona, history, out, tokenizer, with_eos=False, max_len=model.cfg.n_positions)
        ids = torch.tensor(inst.input_ids, device=device).unsqueeze(0)
        tt = torch.tensor(inst.token_type_ids, device=device).unsqueeze(0)
        logits = _lm_logits(model, model.hidden(ids, tt)[0, -1]) / max(temperature, 1e-6)
        probs = F.softmax(top_filtering(logits, top_k, top_p), dim=-1)
        nxt = int(torch.argmax(probs)) if no_sample else int(torch.multinomial(probs, 1, generator=generator))
        if step < min_length and nxt in special:
            tries = 0
            while nxt in special and float(probs.max()) < 1.0 and tries < 64:
                nxt = int(torch.multinomial(probs, 1, generator=generator))
                tries += 1
        if nxt in special:
            break
        out.append(nxt)
    return out


# ----------------------------------------------------------------------------------------------------------------
# metrics
# ----------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def evaluate_tensors(model, loader: Iterable[Dict[str, torch.Tensor]], device="cpu") -> Dict[str, float]:
    """hits@1 (multiple-choice accuracy), token NLL / perplexity of the gold replies over a validation loader of
    ``[B, C, T]`` batches (``data.get_data_loaders``); the same quantities the reference's training script tracks as
    ``accuracy`` / ``nll`` / ``average_ppl`` (train_gpt2_ddp.py:201-208)."""
    hits = total = 0
    nll_sum, n_tok = 0.0, 0
    for batch in loader:
        b = {k: v.to(device) for k, v in batch.items()}
        B, C, T = b["input_ids"].shape
        h = model.hidden(b["input_ids"].reshape(B * C, T), b["token_type_ids"].reshape(B * C, T))
        idx = b["mc_token_ids"].reshape(B * C, 1, 1).expand(-1, 1, h.shape[-1])
        mc = model.mc_head(h.gather(1, idx).squeeze(1)).view(B, C).float()
        hits += int((mc.argmax(-1) == b["mc_labels"]).sum())
        total += B
        gold = h.view(B, C, T, -1)[:, -1]                                   # LM labels live on the last candidate
        labels = b["lm_labels"][:, -1, 1:].reshape(-1)
        keep = labels != IGNORE
        if keep.any():
            rows = gold[:, :-1].reshape(-1, gold.shape[-1])[keep]
            nll_sum += float(F.cross_entropy(_lm_logits(model, rows), labels[keep], reduction="sum"))
            n_tok += int(keep.sum())
    nll = nll_sum / max(1, n_tok)
    return {"hits@1": hits / max(1, total), "nll": nll, "ppl": math.exp(min(nll, 50.0)), "examples": total, "tokens": n_tok,
            "hits": hits, "nll_sum": nll_sum}


_ARTICLES = re.compile(r"\b(a|an|the)\b")
_PUNCT = str.maketrans({c: " " for c in string.punctuation})


def normalize_answer(s: str) -> List[str]:
    return _ARTICLES.sub(" ", s.lower().translate(_PUNCT)).split()


def f1_score(guess: str, answers: Sequence[str]) -> float:
    """Best word-overlap F1 of ``guess`` against any of ``answers``."""
    g = normalize_answer(guess)
    best = 0.0
    for a in answers:
        ref = normalize_answer(a)
        common = sum((Counter(g) & Counter(ref)).values())
        if common:
            p, r = common / len(g), common / len(ref)
            best = max(best, 2 * p * r / (p + r))
    return best


def build_prefix2words(tokenizer: DialogTokenizer, word_freq: Dict[str, int], smoothing: int = 5) -> Dict[int, Dict[str, float]]:
    """first BPE token of a word → {word: share of that prefix's (smoothed) frequency mass}."""
    table: Dict[int, Dict[str, float]] = defaultdict(dict)
    for word, freq in word_freq.items():
        table[tokenizer.first_symbol_id(" " + word)][word] = freq + smoothing
    for words in table.values():
        z = sum(words.values())
        for w in words:
            words[w] /= z
    return dict(table)


@torch.no_grad()
def next_word_probability(model, tokenizer: DialogTokenizer, persona, history, partial_out: Sequence[str],
                          prefix2words: Dict[int, Dict[str, float]], device="cpu") -> Dict[str, float]:
    """Distribution over dictionary words for the next word of the reply, given the words produced so far."""
    partial = tokenizer.encode(" " + " ".join(partial_out)) if partial_out else []
    inst = build_input_from_segments(persona, history, partial, tokenizer, with_eos=False, max_len=model.cfg.n_positions)
    ids = torch.tensor(inst.input_ids, device=device).unsqueeze(0)
    tt = torch.tensor(inst.token_type_ids, device=device).unsqueeze(0)
    probs = F.softmax(_lm_logits(model, model.hidden(ids, tt)[0, -1]), dim=-1)
    dist: Dict[str, float] = {}
    for prefix, words in prefix2words.items():
        p = float(probs[prefix])
        for w, share in words.items():
            dist[w] = p * share
    return dist


@torch.no_grad()
def evaluate_dialogs(model, tokenizer: DialogTokenizer, raw_split: list, eval_type: str = "f1", max_history: int = 2,
                     max_examples: int = 0, device="cpu", seed: int = 0, **sample_kw) -> Dict[str, float]:
    """Run one metric over a RAW (string) PersonaChat split the way an agent would see it: persona + running history,
    gold reply = last candidate. ``eval_type``: ``hits@1`` | ``ppl`` (word-level) | ``f1``."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    word_freq: Counter = Counter()
    if eval_type == "ppl":
        for d in raw_split:
            for u in d["utterances"]:
                word_freq.update(normalize_answer(u["candidates"][-1]))
        table = build_prefix2words(tokenizer, word_freq)
    n = 0
    acc = 0.0
    logloss, words = 0.0, 0
    for d in raw_split:
        persona = [tokenizer.encode(p) for p in d["personality"]]
        for u in d["utterances"]:
            if max_examples and n >= max_examples:
                break
            hist = [tokenizer.encode(h) for h in u["history"][-(2 * max_history + 1):]]
            gold = u["candidates"][-1]
            if eval_type == "hits@1":
                insts = [build_input_from_segments(persona, hist, tokenizer.encode(c), tokenizer, max_len=model.cfg.n_positions)
                         for c in u["candidates"]]
                T = max(len(i.input_ids) for i in insts)
                pad = tokenizer.pad_id
                ids = torch.tensor([i.input_ids + [pad] * (T - len(i.input_ids)) for i in insts], device=device)[None]
                tt = torch.tensor([i.token_type_ids + [pad] * (T - len(i.token_type_ids)) for i in insts], device=device)[None]
                mc = torch.tensor([i.mc_token_id for i in insts], device=device)[None]
                acc += float(int(mc_scores(model, ids, tt, mc)[0].argmax()) == len(insts) - 1)
            elif eval_type == "ppl":
                said: List[str] = []
                for w in normalize_answer(gold):
                    dist = next_word_probability(model, tokenizer, persona, hist, said, table, device)
                    z = sum(dist.values()) or 1.0
                    logloss -= math.log(max(dist.get(w, 0.0) / z, 1e-7))
                    words += 1
                    said.append(w)
            else:
                out = sample_reply(model, tokenizer, persona, hist, device=device, generator=gen, **sample_kw)
                acc += f1_score(tokenizer.decode(out), [gold])
            n += 1
    if eval_type == "ppl":
        return {"ppl": math.exp(min(logloss / max(1, words), 50.0)), "words": words, "examples": n}
    return {eval_type: acc / max(1, n), "examples": n}
