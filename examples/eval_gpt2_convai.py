"""ConvAI2-style evaluation of a fine-tuned dialogue model: ``--eval_type hits@1 | ppl | f1``
(the reference's /root/reference/models/gpt2/convai_evaluation.py, which drives ParlAI's eval_hits / eval_ppl / eval_f1;
here the three metrics are computed directly, see adapcc_b200/eval/convai.py).

    python examples/eval_gpt2_convai.py --model_checkpoint ck.pt --eval_type hits@1 [--dataset_path personachat.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adapcc_b200.data import synthetic_personachat  # noqa: E402
from adapcc_b200.eval import evaluate_dialogs, load_model_from_checkpoint  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--model_checkpoint", required=True, help="file written by examples/train_gpt2_personachat.py --checkpoint")
    p.add_argument("--dataset_path", default="", help="PersonaChat-schema JSON (its 'valid' split is scored); empty: synthetic")
    p.add_argument("--synthetic_dialogs", type=int, default=256)
    p.add_argument("--eval_type", default="hits@1", choices=["hits@1", "ppl", "f1"])
    p.add_argument("--max_history", type=int, default=2)
    p.add_argument("--max_examples", type=int, default=0)
    p.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    p.add_argument("--no_sample", action="store_true")
    p.add_argument("--max_length", type=int, default=20)
    p.add_argument("--min_length", type=int, default=1)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--temperature", type=float, default=0.7)
    p.add_argument("--top_k", type=int, default=20)
    p.add_argument("--top_p", type=float, default=0.9)
    a = p.parse_args()
    model, tok = load_model_from_checkpoint(a.model_checkpoint, a.device,
                                            torch.bfloat16 if a.device.startswith("cuda") else None)
    if a.dataset_path:
        raw = json.load(open(a.dataset_path, encoding="utf-8"))
    else:
        raw = synthetic_personachat(n_train=a.synthetic_dialogs, n_valid=max(8, a.synthetic_dialogs // 8), n_candidates=4, seed=a.seed)
    kw = {}
    if a.eval_type == "f1":
        kw = dict(max_length=a.max_length, min_length=a.min_length, temperature=a.temperature, top_k=a.top_k, top_p=a.top_p,
                  no_sample=a.no_sample)
    res = evaluate_dialogs(model, tok, raw["valid"], a.eval_type, a.max_history, a.max_examples, a.device, a.seed, **kw)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
