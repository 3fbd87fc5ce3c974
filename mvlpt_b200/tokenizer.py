"""Tokenizer plumbing (init-time only; SURVEY.md §2 #2-3 marks BPE itself out of scope for this build).

Class-name prompts are tokenised once, when the prompt learner is constructed (trainers/mvlpt.py:292-305).  The BPE
vocabulary ships only with the reference (clip/bpe_simple_vocab_16e6.txt.gz), so:
  * if the reference's `clip` package is importable (it is wherever this drops into an MVLPT checkout) its
    SimpleTokenizer is used unchanged;
  * callers may pass `tokenized_prompts` / `name_lens` directly (the golden fixtures do);
  * benchmarks use `SyntheticTokenizer` — "random class-name token sequences" as BASELINE.json asks.
"""
from __future__ import annotations

import zlib
from typing import List, Optional

import torch

SOT, EOT = 49406, 49407
_override = None


class SyntheticTokenizer:
    """Deterministic stand-in: one token per whitespace-separated word ('.' split off), ids hashed into the BPE range.
    'X' maps to 343 like the real vocabulary (SURVEY.md §8c KATs) so placeholder rows look the same."""

    encoder = {"<|startoftext|>": SOT, "<|endoftext|>": EOT}

    def encode(self, text: str) -> List[int]:
        out = []
        for w in text.lower().replace(".", " . ").split():
            if w == "x":
                out.append(343)
            elif w == ".":
                out.append(269)
            else:
                out.append(1000 + zlib.crc32(w.encode()) % 47000)
        return out


def set_tokenizer(tok) -> None:
    global _override
    _override = tok


def get_tokenizer():
    if _override is not None:
        return _override
    try:
        from clip.simple_tokenizer import SimpleTokenizer  # the reference's own module, if on sys.path
        return SimpleTokenizer()
    except Exception as exc:  # pragma: no cover - depends on the deployment
        raise RuntimeError(
            "no BPE tokenizer available: put the MVLPT checkout (its `clip` package) on sys.path, call "
            "mvlpt_b200.tokenizer.set_tokenizer(...), or pass tokenized_prompts/name_lens explicitly") from exc


def tokenize(texts, context_length: int = 77, tokenizer=None) -> torch.Tensor:
    """Same contract as clip.tokenize (clip/clip.py:187-223): [SOT] + bpe + [EOT], zero padded, error if too long."""
    tok = tokenizer or get_tokenizer()
    if isinstance(texts, str):
        texts = [texts]
    out = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, t in enumerate(texts):
        ids = [SOT] + tok.encode(t) + [EOT]
        if len(ids) > context_length:
            raise RuntimeError(f"Input {t} is too long for context length {context_length}")
        out[i, :len(ids)] = torch.tensor(ids)
    return out
