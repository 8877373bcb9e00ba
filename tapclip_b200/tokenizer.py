"""Tokenizer contract of the hot path: ``tokenizer(str) -> LongTensor[1, 77]`` (SOT 49406, EOT 49407, pad 0).

The reference takes ``open_clip.get_tokenizer(model_name)`` (models/clip_wrapper.py:27); its BPE vocabulary
ships inside open_clip, which is not available offline.  Any callable with the same contract can be
passed to ``CLIPWrapper(tokenizer=...)``; this deterministic stand-in is the default so that
``PromptLearner.add_class_prompt`` (models/prompt_learner.py:31-34) works without that dependency.
Tokenisation is one-off setup, not part of the hot path.
"""
from __future__ import annotations

import zlib

import torch


class SyntheticTokenizer:
    SOT, EOT = 49406, 49407
    COMMON = {"a": 320, "photo": 1125, "of": 539}

    def __init__(self, context_length: int = 77):
        self.context_length = context_length

    def encode_word(self, word: str):
        if word in self.COMMON:
            return [self.COMMON[word]]
        n_pieces = 1 + zlib.crc32(word.encode("utf-8")) % 3
        return [1000 + zlib.crc32(f"{word}#{i}".encode("utf-8")) % 48405 for i in range(n_pieces)]

    def __call__(self, texts, context_length: int | None = None):
        texts = [texts] if isinstance(texts, str) else list(texts)
        length = context_length or self.context_length
        ids = torch.zeros(len(texts), length, dtype=torch.long)
        for row, text in enumerate(texts):
            seq = [self.SOT]
            for word in text.lower().split():
                seq += self.encode_word(word)
            seq = seq[: length - 1] + [self.EOT]
            ids[row, : len(seq)] = torch.tensor(seq)
        return ids
