"""Chat with a fine-tuned dialogue model: a persona is drawn from the dataset, every user line is appended to the
history and answered by nucleus / top-k sampling (the reference's /root/reference/models/gpt2/interact.py:90-150).

    python examples/interact_gpt2.py --model_checkpoint ck.pt            # interactive
    python examples/interact_gpt2.py --model_checkpoint ck.pt --script "hi there|what do you do for fun ?"
"""
import argparse
import json
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adapcc_b200.data import synthetic_personachat  # noqa: E402
from adapcc_b200.eval import load_model_from_checkpoint, sample_reply  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--model_checkpoint", required=True)
    p.add_argument("--dataset_path", default="", help="personas are drawn from this PersonaChat-schema JSON; empty: synthetic")
    p.add_argument("--max_history", type=int, default=2)
    p.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    p.add_argument("--no_sample", action="store_true")
    p.add_argument("--max_length", type=int, default=20)
    p.add_argument("--min_length", type=int, default=1)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--temperature", type=float, default=0.7)
    p.add_argument("--top_k", type=int, default=0)
    p.add_argument("--top_p", type=float, default=0.9)
    p.add_argument("--script", default="", help="'|'-separated user lines instead of stdin (non-interactive runs)")
    a = p.parse_args()
    random.seed(a.seed)
    gen = torch.Generator().manual_seed(a.seed)
    model, tok = load_model_from_checkpoint(a.model_checkpoint, a.device,
                                            torch.bfloat16 if a.device.startswith("cuda") else None)
    raw = json.load(open(a.dataset_path, encoding="utf-8")) if a.dataset_path else synthetic_personachat(seed=a.seed)
    persona_text = random.choice([d["personality"] for split in raw.values() for d in split])
    print("Selected personality:", " ".join(persona_text))
    persona = [tok.encode(s) for s in persona_text]
    history = []
    lines = iter(a.script.split("|")) if a.script else None
    while True:
        try:
            text = next(lines) if lines is not None else input(">>> ")
        except (StopIteration, EOFError):
            break
        if not text.strip():
            print("Prompt should not be empty!")
            continue
        history.append(tok.encode(text))
        out = sample_reply(model, tok, persona, history, a.max_length, a.min_length, a.temperature, a.top_k, a.top_p,
                           a.no_sample, a.device, gen)
        history.append(out)
        history = history[-(2 * a.max_history + 1):]
        print(tok.decode(out))


if __name__ == "__main__":
    main()
