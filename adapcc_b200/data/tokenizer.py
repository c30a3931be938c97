"""Self-contained byte-level BPE tokenizer for the conversational GPT-2 workload.

The reference tokenizes PersonaChat with HuggingFace's pretrained ``GPT2Tokenizer`` (downloaded vocabulary,
/root/reference/models/gpt2/train_gpt2_ddp.py:150-153) and adds five special tokens
(``<bos> <eos> <speaker1> <speaker2> <pad>``, train_gpt2_ddp.py:28-31). There is no network here, so the vocabulary
is *trained* from the corpus at hand (byte-level BPE: 256 byte symbols + learned merges, GPT-2's scheme) and stored
as one JSON file; a directory with ``vocab.json`` / ``merges.txt`` of the real GPT-2 vocabulary is used instead when
it is given (``DialogTokenizer.from_gpt2_files``). Either way the five special tokens take the TOP five ids of the
model's vocabulary (``GPT2Config.vocab_size`` = 50257 + 5), the layout the reference's ``add_special_tokens_`` produces.
"""
from __future__ import annotations

import json
import os
import re
from collections import Counter, defaultdict
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

SPECIAL_TOKENS = ("<bos>", "<eos>", "<speaker1>", "<speaker2>", "<pad>")

# words keep their leading blank (GPT-2's convention), digits and punctuation are split off
# (the underscore is a "word" character for `re` but neither a letter nor a digit: it belongs to the punctuation class)
_PRETOKEN = re.compile(r"'s|'t|'re|'ve|'m|'ll|'d| ?[^\W\d_]+| ?\d+| ?(?:[^\s\w]|_)+|\s+(?!\S)|\s+", re.UNICODE)


def _byte_symbols() -> Tuple[Dict[int, str], Dict[str, int]]:
    """Printable stand-ins for the 256 byte values (so a merge table is a plain text file)."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    table, nxt = {}, 0
    for b in range(256):
        if b in keep:
            table[b] = chr(b)
        else:
            table[b] = chr(256 + nxt)
            nxt += 1
    return table, {c: b for b, c in table.items()}


_B2S, _S2B = _byte_symbols()


class DialogTokenizer:
    """``encode(text) -> ids``, ``decode(ids) -> text``; ``special_ids`` = (bos, eos, speaker1, speaker2, pad)."""

    def __init__(self, symbols: Sequence[str], merges: Sequence[Tuple[str, str]], model_vocab: Optional[int] = None):
        self.symbols = list(symbols)
        self.index = {s: i for i, s in enumerate(self.symbols)}
        self.merges = [tuple(m) for m in merges]
        self.rank = {m: i for i, m in enumerate(self.merges)}
        self.base_vocab = len(self.symbols)
        self.model_vocab = int(model_vocab) if model_vocab else self.base_vocab + len(SPECIAL_TOKENS)
        if self.model_vocab < self.base_vocab + len(SPECIAL_TOKENS):
            raise ValueError("model vocabulary smaller than the tokenizer's")
        first = self.model_vocab - len(SPECIAL_TOKENS)
        self.special = {t: first + i for i, t in enumerate(SPECIAL_TOKENS)}
        self._cache: Dict[str, List[int]] = {}

    # ------------------------------------------------------------------ properties
    @property
    def vocab_size(self) -> int:
        return self.model_vocab

    @property
    def special_ids(self) -> Tuple[int, int, int, int, int]:
        return tuple(self.special[t] for t in SPECIAL_TOKENS)

    @property
    def pad_id(self) -> int:
        return self.special["<pad>"]

    def convert_tokens_to_ids(self, tokens):
        """Reference-style helper (``tokenizer.convert_tokens_to_ids(SPECIAL_TOKENS)``)."""
        if isinstance(tokens, str):
            return self.special.get(tokens, self.index.get(tokens))
        return [self.convert_tokens_to_ids(t) for t in tokens]

    def fingerprint(self) -> str:
        """Short hash of the vocabulary (cache keys: a tokenized dataset is only valid for the tokenizer that made it)."""
        import hashlib

        h = hashlib.sha1()
        for s in self.symbols:
            h.update(s.encode("utf-8") + b"\0")
        h.update(str(self.model_vocab).encode())
        return h.hexdigest()[:12]

    # ------------------------------------------------------------------ training
    @classmethod
    def train(cls, corpus: Iterable[str], vocab_size: int = 2048, model_vocab: Optional[int] = None,
              min_pair_count: int = 2) -> "DialogTokenizer":
        """Learn ``vocab_size - 256`` merges from ``corpus`` (word-frequency table + pair index: each merge touches only
        the words that contain the pair)."""
        words = Counter()
        for text in corpus:
            for w in _PRETOKEN.findall(text):
                words["".join(_B2S[b] for b in w.encode("utf-8"))] += 1
        seqs = {w: list(w) for w in words}
        pairs: Counter = Counter()
        where: Dict[Tuple[str, str], set] = defaultdict(set)
        for w, seq in seqs.items():
            for a, b in zip(seq, seq[1:]):
                pairs[(a, b)] += words[w]
                where[(a, b)].add(w)
        symbols = [_B2S[b] for b in range(256)]
        merges: List[Tuple[str, str]] = []
        while len(symbols) < vocab_size and pairs:
            (a, b), cnt = max(pairs.items(), key=lambda kv: (kv[1], kv[0]))
            if cnt < min_pair_count:
                break
            merges.append((a, b))
            symbols.append(a + b)
            for w in list(where.pop((a, b), ())):
                seq, f = seqs[w], words[w]
                for x, y in zip(seq, seq[1:]):                    # retire the word's old pairs
                    pairs[(x, y)] -= f
                    if pairs[(x, y)] <= 0:
                        pairs.pop((x, y), None)
                    where[(x, y)].discard(w)
                out, i = [], 0
                while i < len(seq):
                    if i + 1 < len(seq) and seq[i] == a and seq[i + 1] == b:
                        out.append(a + b)
                        i += 2
                    else:
                        out.append(seq[i])
                        i += 1
                seqs[w] = out
                for x, y in zip(out, out[1:]):
                    pairs[(x, y)] += f
                    where[(x, y)].add(w)
            pairs.pop((a, b), None)
        return cls(symbols, merges, model_vocab)

    # ------------------------------------------------------------------ persistence
    def save(self, path: str) -> None:
        tmp = path + ".tmp"
        with open(tmp, "w", encoding="utf-8") as f:
            json.dump({"symbols": self.symbols, "merges": self.merges, "model_vocab": self.model_vocab}, f)
        os.replace(tmp, path)

    @classmethod
    def load(cls, path: str) -> "DialogTokenizer":
        with open(path, encoding="utf-8") as f:
            d = json.load(f)
        return cls(d["symbols"], [tuple(m) for m in d["merges"]], d.get("model_vocab"))

    @classmethod
    def from_gpt2_files(cls, directory: str, model_vocab: Optional[int] = None) -> "DialogTokenizer":
        """A directory holding the published GPT-2 ``vocab.json`` + ``merges.txt`` (same byte-symbol alphabet)."""
        with open(os.path.join(directory, "vocab.json"), encoding="utf-8") as f:
            vocab = json.load(f)
        symbols = [s for s, _ in sorted(vocab.items(), key=lambda kv: kv[1])]
        merges = []
        with open(os.path.join(directory, "merges.txt"), encoding="utf-8") as f:
            for line in f:
                parts = line.rstrip("\n").split(" ")
                if len(parts) == 2 and not line.startswith("#"):
                    merges.append((parts[0], parts[1]))
        return cls(symbols, merges, model_vocab or len(symbols) + len(SPECIAL_TOKENS))

    # ------------------------------------------------------------------ encode / decode
    def _bpe(self, word: str) -> List[int]:
        got = self._cache.get(word)
        if got is not None:
            return got
        seq = list(word)
        while len(seq) > 1:
            best, at = None, -1
            for i, pair in enumerate(zip(seq, seq[1:])):
                r = self.rank.get(pair)
                if r is not None and (best is None or r < best):
                    best, at = r, i
            if best is None:
                break
            a, b = seq[at], seq[at + 1]
            out, i = [], 0
            while i < len(seq):
                if i + 1 < len(seq) and seq[i] == a and seq[i + 1] == b:
                    out.append(a + b)
                    i += 2
                else:
                    out.append(seq[i])
                    i += 1
            seq = out
        ids = [self.index[s] for s in seq]
        if len(self._cache) < 200_000:
            self._cache[word] = ids
        return ids

    def encode(self, text: str) -> List[int]:
        ids: List[int] = []
        for w in _PRETOKEN.findall(text):
            ids.extend(self._bpe("".join(_B2S[b] for b in w.encode("utf-8"))))
        return ids

    def decode(self, ids: Iterable[int], skip_special_tokens: bool = True) -> str:
        names = {v: k for k, v in self.special.items()}
        out, buf = [], bytearray()
        for i in ids:
            i = int(i)
            if i in names:
                if not skip_special_tokens:
                    out.append(buf.decode("utf-8", errors="replace"))
                    buf = bytearray()
                    out.append(names[i])
                continue
            if 0 <= i < self.base_vocab:
                buf.extend(_S2B[c] for c in self.symbols[i])
        out.append(buf.decode("utf-8", errors="replace"))
        return "".join(out)

    def first_symbol_id(self, word: str) -> int:
        """Id of the first BPE token of ``word`` (the prefix table of the per-word perplexity evaluation,
        /root/reference/models/gpt2/convai_evaluation.py:182-197)."""
        return self.encode(word)[0]
