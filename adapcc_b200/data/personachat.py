"""PersonaChat-format dialogue data for the GPT-2 double-heads workload: loading / tokenizing / caching, the
persona + history + reply input layout, static-shape tensors and (distributed) loaders.

What the reference does (/root/reference/models/gpt2/utils.py:34-56, train_gpt2_ddp.py:47-118): download
``personachat_self_original.json``, tokenize every string, build for every utterance ``num_candidates`` sequences
``<bos> persona  <speaker?> history…  <speaker?> reply <eos>`` with token types = speaker of the segment, score the LM
loss only on the gold (last) candidate's reply and the multiple-choice loss on the position of the last token, pad to
the longest sequence, wrap in a ``TensorDataset`` + ``DistributedSampler``.

Here: the same JSON schema is read from a local file when there is one and *generated* otherwise (no network:
``synthetic_personachat`` writes template dialogues whose gold reply depends on the persona, so both heads have
something to learn); sequences are laid out in ONE pass into preallocated ``[N, C, T]`` int64 arrays with a fixed
``T`` (static shapes → the training step stays CUDA-graph replayable, rows stay 64-aligned for the fused LM-head
kernels) and over-long inputs lose their OLDEST history first instead of overflowing the position table.
"""
from __future__ import annotations

import json
import os
import random
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

from .tokenizer import DialogTokenizer

MODEL_INPUTS = ("input_ids", "mc_token_ids", "lm_labels", "mc_labels", "token_type_ids")
IGNORE = -100

# ----------------------------------------------------------------------------------------------------------------
# synthetic corpus in the PersonaChat schema
# ----------------------------------------------------------------------------------------------------------------
_FACTS = {
    "hobby": ("i like {} .", "what do you do for fun ?", "i spend my free time {} .",
              ["hiking", "painting", "chess", "fishing", "knitting", "surfing", "baking", "gardening", "running", "gaming"]),
    "food": ("my favorite food is {} .", "what do you like to eat ?", "i could eat {} every day .",
             ["pizza", "sushi", "tacos", "curry", "pasta", "ramen", "salad", "burgers", "dumplings", "waffles"]),
    "pet": ("i have a {} .", "do you have any pets ?", "yes , my {} keeps me company .",
            ["dog", "cat", "parrot", "hamster", "turtle", "rabbit", "goldfish", "lizard", "horse", "ferret"]),
    "job": ("i work as a {} .", "what do you do for a living ?", "i am a {} , it pays the bills .",
            ["teacher", "nurse", "pilot", "chef", "plumber", "lawyer", "farmer", "coder", "driver", "painter"]),
    "city": ("i live in {} .", "where are you from ?", "i am from {} , it is lovely there .",
             ["paris", "tokyo", "austin", "lima", "oslo", "cairo", "delhi", "sydney", "berlin", "seoul"]),
}
_OPENERS = ["hi , how are you today ?", "hello there !", "hey , nice to meet you .", "good evening , how is it going ?"]
_SMALLTALK = ["i am doing well , thanks .", "pretty good , just relaxing .", "not bad at all .", "great , thank you for asking ."]


def synthetic_personachat(n_train: int = 64, n_valid: int = 16, n_candidates: int = 4, turns: int = 4,
                          seed: int = 0) -> Dict[str, list]:
    """Template dialogues in the schema of ``personachat_self_original.json``:
    ``{"train": [{"personality": [str], "utterances": [{"history": [str], "candidates": [str]}]}], "valid": [...]}``
    — the gold reply is the LAST candidate (the convention the reference relies on, train_gpt2_ddp.py:93-99)."""
    rng = random.Random(seed)
    keys = list(_FACTS)

    def one_dialog():
        picks = rng.sample(keys, 4)
        values = {k: rng.choice(_FACTS[k][3]) for k in picks}
        persona = [_FACTS[k][0].format(values[k]) for k in picks]
        history, utterances = [], []
        asked = picks[:]
        rng.shuffle(asked)
        for t in range(turns):
            if t == 0:
                question, gold = rng.choice(_OPENERS), rng.choice(_SMALLTALK)
            else:
                k = asked[(t - 1) % len(asked)]
                question, gold = _FACTS[k][1], _FACTS[k][2].format(values[k])
            history = history + [question]
            distractors = []
            while len(distractors) < n_candidates - 1:
                k2 = rng.choice(keys)
                cand = _FACTS[k2][2].format(rng.choice(_FACTS[k2][3])) if rng.random() < 0.8 else rng.choice(_SMALLTALK)
                if cand != gold and cand not in distractors:
                    distractors.append(cand)
            utterances.append({"history": list(history), "candidates": distractors + [gold]})
            history = history + [gold]
        return {"personality": persona, "utterances": utterances}

    return {"train": [one_dialog() for _ in range(n_train)], "valid": [one_dialog() for _ in range(n_valid)]}


def corpus_of(dataset: Dict[str, list]) -> Iterable[str]:
    """Every string of a PersonaChat-schema dataset (tokenizer training)."""
    for split in dataset.values():
        for dialog in split:
            yield from dialog["personality"]
            for u in dialog["utterances"]:
                yield from u["history"]
                yield from u["candidates"]


# ----------------------------------------------------------------------------------------------------------------
# load / tokenize / cache
# ----------------------------------------------------------------------------------------------------------------
def _tokenize(obj, tok: DialogTokenizer):
    if isinstance(obj, str):
        return tok.encode(obj)
    if isinstance(obj, dict):
        return {k: _tokenize(v, tok) for k, v in obj.items()}
    return [_tokenize(v, tok) for v in obj]


def get_dataset(tokenizer: DialogTokenizer, dataset_path: str = "", dataset_cache: str = "",
                synthetic: Optional[dict] = None) -> Dict[str, list]:
    """Tokenized dataset: from ``dataset_cache`` if it exists, else from the JSON file at ``dataset_path``, else a
    synthetic one (``synthetic`` = kwargs of :func:`synthetic_personachat`). Counterpart of the reference's
    ``get_dataset`` (utils.py:34-56) minus the S3 download."""
    cache = f"{dataset_cache}_{type(tokenizer).__name__}_{tokenizer.fingerprint()}" if dataset_cache else ""
    if cache and os.path.isfile(cache):
        return torch.load(cache, weights_only=False)
    if dataset_path:
        with open(dataset_path, encoding="utf-8") as f:
            raw = json.load(f)
    else:
        raw = synthetic_personachat(**(synthetic or {}))
    data = _tokenize(raw, tokenizer)
    if cache:
        tmp = f"{cache}.{os.getpid()}.tmp"              # every rank may get here at once: private file, atomic publish
        torch.save(data, tmp)
        os.replace(tmp, cache)
    return data


# ----------------------------------------------------------------------------------------------------------------
# input layout
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class Instance:
    input_ids: List[int]
    token_type_ids: List[int]
    lm_labels: List[int]
    mc_token_id: int


def build_input_from_segments(persona: Sequence[Sequence[int]], history: Sequence[Sequence[int]], reply: Sequence[int],
                              tokenizer: DialogTokenizer, lm_labels: bool = False, with_eos: bool = True,
                              max_len: Optional[int] = None) -> Instance:
    """One model input from persona sentences, dialogue history and a (candidate) reply. Layout and labels as in the
    reference (train_gpt2_ddp.py:61-74): segment 0 = ``<bos>`` + all persona tokens; every later segment starts with a
    speaker token, alternating so that the reply is spoken by ``<speaker2>``'s counterpart of the last history turn;
    token types alternate per segment starting with ``<speaker1>`` for the persona; only the reply's tokens (after its
    speaker token) carry LM labels. With ``max_len`` the oldest history turns are dropped until the input fits (the
    reference has no bound and would index past the position table)."""
    bos, eos, sp1, sp2, _ = tokenizer.special_ids
    history = [list(h) for h in history]
    while True:
        segments: List[List[int]] = [[bos] + [t for sent in persona for t in sent]]
        turns = history + [list(reply) + ([eos] if with_eos else [])]
        n = len(turns) + 1
        for i, turn in enumerate(turns):
            segments.append([sp2 if (n - i) % 2 else sp1] + turn)
        total = sum(len(s) for s in segments)
        if max_len is None or total <= max_len or not history:
            break
        history = history[1:]
    ids = [t for s in segments for t in s]
    types = [(sp2 if k % 2 else sp1) for k, s in enumerate(segments) for _ in s]
    labels = [IGNORE] * len(ids)
    if lm_labels:
        start = len(ids) - len(segments[-1]) + 1               # first token after the reply's speaker token
        labels[start:] = segments[-1][1:]
    if max_len is not None and len(ids) > max_len:             # a persona + reply longer than the window: keep the tail
        cut = len(ids) - max_len
        ids, types, labels = ids[cut:], types[cut:], labels[cut:]
    return Instance(ids, types, labels, len(ids) - 1)


def build_tensors(split: list, tokenizer: DialogTokenizer, num_candidates: int = 2, max_history: int = 2,
                  personality_permutations: int = 1, seq_len: Optional[int] = None, align: int = 64,
                  limit_candidates: bool = True) -> Dict[str, torch.Tensor]:
    """``[N, C, T]`` tensors for one split (N = utterances × permutations). ``seq_len=None`` pads to the longest input
    rounded up to ``align``. ``limit_candidates=False`` keeps every candidate (validation: the reference only trims the
    training split, train_gpt2_ddp.py:84-86)."""
    have = len(split[0]["utterances"][0]["candidates"])
    C = min(num_candidates, have) if (limit_candidates and num_candidates > 0) else have
    rows: List[List[Instance]] = []
    for dialog in split:
        persona = [list(p) for p in dialog["personality"]]
        for _ in range(personality_permutations):
            for utt in dialog["utterances"]:
                hist = utt["history"][-(2 * max_history + 1):]
                cands = utt["candidates"][-C:]
                rows.append([build_input_from_segments(persona, hist, c, tokenizer, lm_labels=(j == C - 1), max_len=seq_len)
                             for j, c in enumerate(cands)])
            persona = [persona[-1]] + persona[:-1]
    longest = max(len(inst.input_ids) for r in rows for inst in r)
    T = seq_len or (longest + align - 1) // align * align
    N = len(rows)
    ids = np.full((N, C, T), tokenizer.pad_id, dtype=np.int64)
    types = np.full((N, C, T), tokenizer.pad_id, dtype=np.int64)
    labels = np.full((N, C, T), IGNORE, dtype=np.int64)
    mc_tok = np.zeros((N, C), dtype=np.int64)
    for n, r in enumerate(rows):
        for c, inst in enumerate(r):
            L = len(inst.input_ids)
            ids[n, c, :L] = inst.input_ids
            types[n, c, :L] = inst.token_type_ids
            labels[n, c, :L] = inst.lm_labels
            mc_tok[n, c] = inst.mc_token_id
    return {"input_ids": torch.from_numpy(ids), "mc_token_ids": torch.from_numpy(mc_tok),
            "lm_labels": torch.from_numpy(labels), "mc_labels": torch.full((N,), C - 1, dtype=torch.int64),
            "token_type_ids": torch.from_numpy(types)}


class DialogDataset(torch.utils.data.Dataset):
    """Dict-of-tensors dataset (the reference wraps the same five tensors in a ``TensorDataset``)."""

    def __init__(self, tensors: Dict[str, torch.Tensor]):
        self.tensors = tensors
        self.n = tensors["input_ids"].shape[0]

    def __len__(self) -> int:
        return self.n

    def __getitem__(self, i):
        return {k: v[i] for k, v in self.tensors.items()}


def get_data_loaders(args, tokenizer: DialogTokenizer, distributed: bool = False, rank: int = 0, world_size: int = 1,
                     seq_len: Optional[int] = None):
    """-> (train_loader, valid_loader, train_sampler, valid_sampler). ``args`` fields (defaults = the reference's,
    train_gpt2_ddp.py:123-140): dataset_path, dataset_cache, num_candidates (2), max_history (2),
    personality_permutations (1), train_batch_size (4), valid_batch_size (4). Training batches have a fixed size
    (``drop_last``): the engine's CUDA graph is captured for one shape."""
    g = lambda k, d: getattr(args, k, d)                                                   # noqa: E731
    data = get_dataset(tokenizer, g("dataset_path", ""), g("dataset_cache", ""), g("synthetic", None))
    train = build_tensors(data["train"], tokenizer, g("num_candidates", 2), g("max_history", 2),
                          g("personality_permutations", 1), seq_len)
    valid = build_tensors(data["valid"], tokenizer, g("num_candidates", 2), g("max_history", 2), 1,
                          seq_len, limit_candidates=False)
    tds, vds = DialogDataset(train), DialogDataset(valid)
    ts = vs = None
    if distributed:
        from torch.utils.data.distributed import DistributedSampler

        ts = DistributedSampler(tds, num_replicas=world_size, rank=rank, shuffle=True, drop_last=True)
        vs = DistributedSampler(vds, num_replicas=world_size, rank=rank, shuffle=False)
    tl = torch.utils.data.DataLoader(tds, batch_size=g("train_batch_size", 4), sampler=ts, shuffle=ts is None, drop_last=True)
    vl = torch.utils.data.DataLoader(vds, batch_size=g("valid_batch_size", 4), sampler=vs, shuffle=False)
    return tl, vl, ts, vs
