"""Host-side tokenisers (stay Python, BASELINE north_star) -- same registry / ABC / ids as the
reference's ``twotower/tokenisers.py`` so vocabularies and checkpoints are interchangeable.

Reference contract: ``BaseTokeniser`` ABC twotower/tokenisers.py:10-30; ``CharTokeniser`` :33-110
(sorted unique chars -> ids from 1, unknown char -> 0 == PAD); ``WordTokeniser`` :113-268
(regex ``\\b\\w+\\b``, frequency-ranked ids from 2, PAD=0, UNK=1); ``REGISTRY`` :276,
``build`` :282.  Added here: ``encode_batch`` which writes a whole batch into one contiguous
(pinned) int64 array -- the input format the fused kernels consume.
"""
from __future__ import annotations

import pickle
import re
from abc import ABC, abstractmethod
from collections import Counter
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch


class BaseTokeniser(ABC):
    @abstractmethod
    def fit(self, texts: Sequence[str]):
        ...

    @abstractmethod
    def encode(self, text: str) -> List[int]:
        ...

    @abstractmethod
    def truncate_and_pad(self, sequence: List[int], max_len: int) -> List[int]:
        ...

    @property
    @abstractmethod
    def vocab_size(self) -> int:
        ...

    # -- batch path (new): [len(texts), max_len] int64, zero padded, optionally pinned ----------
    def encode_batch(self, texts: Sequence[str], max_len: int, pin_memory: bool = False,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
        n = len(texts)
        if out is None:
            out = torch.zeros((n, max_len), dtype=torch.int64, pin_memory=pin_memory)
        else:
            out.zero_()
        arr = out.numpy()
        for i, t in enumerate(texts):
            ids = self.encode(t)[:max_len]
            if ids:
                arr[i, :len(ids)] = ids
        return out


def _pad(sequence: List[int], max_len: int, pad: int) -> List[int]:
    n = len(sequence)
    return sequence + [pad] * (max_len - n) if n < max_len else sequence[:max_len]


class CharTokeniser(BaseTokeniser):
    PAD = 0

    def __init__(self):
        self.string_to_index: Dict[str, int] = {}
        self.index_to_string: Dict[int, str] = {}
        self._lut: Optional[np.ndarray] = None

    def fit(self, texts: Sequence[str]):
        chars = sorted(set().union(*map(set, texts))) if len(texts) else []
        self.string_to_index = {c: i + 1 for i, c in enumerate(chars)}          # 0 = padding
        self.index_to_string = {i: c for c, i in self.string_to_index.items()}
        self._lut = None
        return self

    def encode(self, text: str) -> List[int]:
        get = self.string_to_index.get
        return [get(c, 0) for c in text]

    def decode(self, indices: List[int]) -> str:
        return "".join(self.index_to_string.get(i, "?") for i in indices)

    def truncate_and_pad(self, sequence: List[int], max_len: int) -> List[int]:
        return _pad(sequence, max_len, self.PAD)

    @property
    def vocab_size(self) -> int:
        return len(self.string_to_index) + 1

    def encode_batch(self, texts, max_len, pin_memory=False, out=None):
        # vectorised: code-point lookup table (falls back to the dict for astral code points)
        if self._lut is None:
            hi = max((ord(c) for c in self.string_to_index), default=0)
            lut = np.zeros(min(hi, 0xFFFF) + 2, dtype=np.int64)
            for c, i in self.string_to_index.items():
                if ord(c) < lut.shape[0]:
                    lut[ord(c)] = i
            self._lut = lut
        n = len(texts)
        if out is None:
            out = torch.zeros((n, max_len), dtype=torch.int64, pin_memory=pin_memory)
        else:
            out.zero_()
        arr = out.numpy()
        lut = self._lut
        for i, t in enumerate(texts):
            t = t[:max_len]
            if not t:
                continue
            cp = np.frombuffer(t.encode("utf-32-le"), dtype=np.uint32)
            if cp.max() < lut.shape[0] - 1:
                arr[i, :cp.shape[0]] = lut[cp]
            else:
                arr[i, :len(t)] = self.encode(t)
        return out

    def save(self, filepath: str):
        with open(filepath, "wb") as f:
            pickle.dump(self.string_to_index, f)

    @classmethod
    def load(cls, filepath: str):
        tok = cls()
        with open(filepath, "rb") as f:
            tok.string_to_index = pickle.load(f)
        tok.index_to_string = {i: c for c, i in tok.string_to_index.items()}
        return tok


class WordTokeniser(BaseTokeniser):
    PAD = 0
    UNK = 1

    def __init__(self, lowercase: bool = True, strip_punctuation: bool = True, max_len: int = 32):
        self.word_to_index: Dict[str, int] = {}
        self.index_to_word: Dict[int, str] = {}
        self.lowercase = lowercase
        self.strip_punctuation = strip_punctuation
        self.max_len = max_len
        self._vocab_size = 2
        self.string_to_index = {}
        self.index_to_string = {}
        self.word_pattern = re.compile(r"\b\w+\b")

    def _tokenize(self, text: str) -> List[str]:
        if self.lowercase:
            text = text.lower()
        return self.word_pattern.findall(text) if self.strip_punctuation else text.split()

    def fit(self, texts: Sequence[str]):
        counts: Counter = Counter()
        for t in texts:
            counts.update(self._tokenize(t))
        # frequency-descending, first-seen order among equal counts (stable sort, like the reference)
        ranked = sorted(counts.items(), key=lambda kv: kv[1], reverse=True)
        self.word_to_index = {"<PAD>": self.PAD, "<UNK>": self.UNK}
        for i, (w, _) in enumerate(ranked):
            self.word_to_index[w] = i + 2
        self.index_to_word = {i: w for w, i in self.word_to_index.items()}
        self._vocab_size = len(self.word_to_index)
        self.string_to_index = self.word_to_index
        self.index_to_string = self.index_to_word
        return self

    def encode(self, text: str) -> List[int]:
        get = self.word_to_index.get
        return [get(w, self.UNK) for w in self._tokenize(text)]

    def decode(self, indices: List[int]) -> str:
        return " ".join(self.index_to_word.get(i, "<UNK>") for i in indices if i != self.PAD)

    def truncate_and_pad(self, sequence: List[int], max_len: int = None) -> List[int]:
        return _pad(sequence, self.max_len if max_len is None else max_len, self.PAD)

    @property
    def vocab_size(self) -> int:
        return self._vocab_size

    def save(self, filepath: str):
        with open(filepath, "wb") as f:
            pickle.dump({"word_to_index": self.word_to_index, "lowercase": self.lowercase,
                         "strip_punctuation": self.strip_punctuation, "max_len": self.max_len}, f)

    @classmethod
    def load(cls, filepath: str):
        with open(filepath, "rb") as f:
            data = pickle.load(f)
        tok = cls(lowercase=data.get("lowercase", True), strip_punctuation=data.get("strip_punctuation", True),
                  max_len=data.get("max_len", 32))
        tok.word_to_index = data["word_to_index"]
        tok.index_to_word = {i: w for w, i in tok.word_to_index.items()}
        tok._vocab_size = len(tok.word_to_index)
        tok.string_to_index = tok.word_to_index
        tok.index_to_string = tok.index_to_word
        return tok

    def __call__(self, texts):
        if isinstance(texts, str):
            texts = [texts]
        return torch.tensor([self.truncate_and_pad(self.encode(t)) for t in texts])


REGISTRY = {"char": CharTokeniser, "word": WordTokeniser}


def build(name: str, **kwargs) -> BaseTokeniser:
    if name not in REGISTRY:
        raise ValueError(f"Unknown tokeniser: {name}. Available options: {list(REGISTRY.keys())}")
    return REGISTRY[name](**kwargs)
