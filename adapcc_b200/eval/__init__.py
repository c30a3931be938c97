"""Evaluation of the conversational GPT-2 workload (hits@1 / perplexity / F1) and checkpoint → model loading."""
from __future__ import annotations

from dataclasses import asdict
from typing import Optional, Tuple

import torch

from .convai import (build_prefix2words, evaluate_dialogs, evaluate_tensors, f1_score, mc_scores,  # noqa: F401
                     next_word_probability, normalize_answer, sample_reply, top_filtering)


def pack_checkpoint(model, tokenizer_path: str, engine=None, extra: Optional[dict] = None) -> dict:
    """What ``examples/train_gpt2_personachat.py`` writes: the model configuration, the tokenizer file, and either the
    flat engine's state (fp32 master weights + AdamW moments per parameter name) or a plain ``state_dict``."""
    ck = {"format": "adapcc-gpt2-dialog-1", "config": asdict(model.cfg), "tokenizer": tokenizer_path,
          "engine": engine.state_dict() if engine is not None else None,
          "model": None if engine is not None else {k: v.detach().cpu() for k, v in model.state_dict().items()}}
    ck.update(extra or {})
    return ck


def load_model_from_checkpoint(path: str, device="cpu", dtype: Optional[torch.dtype] = None) -> Tuple[object, object]:
    """-> (model in eval mode, tokenizer). Engine checkpoints are restored from their fp32 master weights."""
    from ..data.tokenizer import DialogTokenizer
    from ..models.gpt2 import GPT2Config, GPT2DoubleHeads

    ck = torch.load(path, map_location="cpu", weights_only=False)
    model = GPT2DoubleHeads(GPT2Config(**ck["config"]))
    if ck.get("engine") is not None:
        per = ck["engine"]["params"]
        with torch.no_grad():
            for name, p in model.named_parameters():
                p.copy_(per[name]["master"])
    else:
        model.load_state_dict(ck["model"])
    model = model.to(device)
    if dtype is not None:
        model = model.to(dtype)
    return model.eval(), DialogTokenizer.load(ck["tokenizer"])
