"""Input pipeline: the reference's ``TripletDataset`` + default collate, restructured for B = 4096 steps.

Reference: ``TripletDataset`` twotower/dataset.py:14-300 -- loads a parquet / tsv file (triplet columns
``query|q_text``, ``positive_doc|d_pos_text``, ``negative_doc|d_neg_text`` :97-122, or ``query, document,
label`` pairs turned into the per-query positive x negative cross product :192-241), fits the tokeniser if it
is empty (:45-48), and keeps one Python list of ints per text; ``__getitem__`` (:262-285) builds three
``torch.tensor(list)`` per item and the DataLoader's default collate stacks 3 x B of them per step.

Here the whole dataset is tokenised ONCE into three contiguous pinned int32 arrays ``[N, max_length]``
(``tokenisers.encode_batch(out=...)``); ``__getitem__`` returns int64 rows like the reference (so the
reference's own DataLoader / train loop work unchanged), and ``batches()`` is the fast path: contiguous
``[B, L]`` slices (sequential order: zero-copy views of the pinned arrays; shuffled: one index_select into a
pinned staging buffer per tensor) that ``FusedTrainer.prefetch`` copies with a single H2D transfer each.
Same constructor, attributes (``query_texts``, ``positive_doc_texts``, ``negative_doc_texts``,
``encoded_queries`` ...), ``vocab_size`` and ``get_original_texts`` as the reference.
"""
from __future__ import annotations

import collections
from typing import Iterator, List, Optional, Sequence, Tuple

import torch
from torch.utils.data import Dataset

from .tokenisers import BaseTokeniser

_QUERY_COLS = ("query", "q_text")
_POS_COLS = ("positive_doc", "d_pos_text")
_NEG_COLS = ("negative_doc", "d_neg_text")


def pairs_to_triplets(queries: Sequence[str], documents: Sequence[str], labels: Sequence[int]):
    """dataset.py:192-241: group by query (first-seen order), keep queries with both kinds, emit pos x neg."""
    groups = collections.defaultdict(lambda: ([], []))
    for q, d, lab in zip(queries, documents, labels):
        groups[q][0 if lab == 1 else 1].append(d)
    out_q, out_p, out_n = [], [], []
    for q, (pos, neg) in groups.items():
        if pos and neg:
            for p in pos:
                for n in neg:
                    out_q.append(q); out_p.append(p); out_n.append(n)
    return out_q, out_p, out_n


class TripletDataset(Dataset):
    def __init__(self, data_path: Optional[str], tokeniser: BaseTokeniser, max_length: int = 64,
                 load_to_memory: bool = True, triplets: Optional[Tuple[Sequence[str], Sequence[str], Sequence[str]]] = None,
                 pin_memory: Optional[bool] = None):
        self.data_path = data_path
        self.tokeniser = tokeniser
        self.max_length = int(max_length)
        self.load_to_memory = load_to_memory
        if triplets is not None:
            q, p, n = triplets
            if not (len(q) == len(p) == len(n)):
                raise ValueError("triplets must be three sequences of equal length")
            self.query_texts, self.positive_doc_texts, self.negative_doc_texts = list(q), list(p), list(n)
        else:
            self._load_data(data_path)
        if not getattr(self.tokeniser, "string_to_index", None):                      # dataset.py:45-48
            self.tokeniser.fit(self.query_texts + self.positive_doc_texts + self.negative_doc_texts)
        self._pin = torch.cuda.is_available() if pin_memory is None else bool(pin_memory)
        self._ids: Optional[List[torch.Tensor]] = None
        self._stage: Optional[List[torch.Tensor]] = None
        if load_to_memory:
            self._tokenise_all()

    # ------------------------------------------------------------------ loading (dataset.py:67-190)
    def _load_data(self, data_path: str) -> None:
        import pandas as pd
        if data_path.endswith(".parquet"):
            df = pd.read_parquet(data_path)
        elif data_path.endswith(".tsv"):
            df = pd.read_csv(data_path, sep="\t")
            if not {"query", "document", "label"} <= set(df.columns):             # headerless synthetic tsv
                df = pd.read_csv(data_path, sep="\t", header=None, names=["query", "document", "label"])
        else:
            raise ValueError(f"Unsupported file format: {data_path}. Supported formats: .parquet, .tsv")
        pick = lambda names: next((c for c in names if c in df.columns), None)
        qc, pc, nc = pick(_QUERY_COLS), pick(_POS_COLS), pick(_NEG_COLS)
        if qc and pc and nc:
            self.query_texts, self.positive_doc_texts, self.negative_doc_texts = \
                df[qc].tolist(), df[pc].tolist(), df[nc].tolist()
        elif {"query", "document", "label"} <= set(df.columns):
            self.query_texts, self.positive_doc_texts, self.negative_doc_texts = pairs_to_triplets(
                df["query"].tolist(), df["document"].tolist(), df["label"].tolist())
        else:
            raise ValueError(f"Unsupported dataframe format with columns: {df.columns.tolist()}. "
                             "Expected either triplets format with columns like 'query'/'q_text', "
                             "'positive_doc'/'d_pos_text', 'negative_doc'/'d_neg_text' "
                             "or pairs format with columns 'query', 'document', 'label'")

    def _tokenise_all(self) -> None:
        n, L = len(self.query_texts), self.max_length
        self._ids = []
        for texts in (self.query_texts, self.positive_doc_texts, self.negative_doc_texts):
            buf = torch.zeros((n, L), dtype=torch.int32, pin_memory=self._pin)
            if hasattr(self.tokeniser, "encode_batch"):
                wide = self.tokeniser.encode_batch(texts, L)
                buf.copy_(wide)
            else:
                for i, t in enumerate(texts):
                    buf[i] = torch.tensor(self.tokeniser.truncate_and_pad(self.tokeniser.encode(t), L), dtype=torch.int32)
            self._ids.append(buf)

    # reference attribute names (lists of int lists in the reference; here [N, L] int32 tensors, same indexing)
    @property
    def encoded_queries(self):
        return self._ids[0]

    @property
    def encoded_positive_docs(self):
        return self._ids[1]

    @property
    def encoded_negative_docs(self):
        return self._ids[2]

    def _encode_and_pad(self, text: str) -> List[int]:
        return self.tokeniser.truncate_and_pad(self.tokeniser.encode(text), self.max_length)

    def __len__(self) -> int:
        return len(self.query_texts)

    def __getitem__(self, index: int):
        """dataset.py:262-285: three int64 [L] tensors."""
        if self._ids is not None:
            return tuple(t[index].long() for t in self._ids)
        return tuple(torch.tensor(self._encode_and_pad(t[index])) for t in
                     (self.query_texts, self.positive_doc_texts, self.negative_doc_texts))

    @property
    def vocab_size(self) -> int:
        return self.tokeniser.vocab_size

    def get_original_texts(self, index: int) -> Tuple[str, str, str]:
        return self.query_texts[index], self.positive_doc_texts[index], self.negative_doc_texts[index]

    # ------------------------------------------------------------------ fast path
    def batches(self, batch_size: int, shuffle: bool = False, drop_last: bool = False,
                generator: Optional[torch.Generator] = None, rank: int = 0, world_size: int = 1
                ) -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        """Yield (q_ids, pos_ids, neg_ids), each a contiguous [b, L] int32 tensor in (pinned) host memory.

        Data parallel: every rank draws the same permutation (pass generators with the same seed) and takes rows
        ``[rank * batch_size, (rank + 1) * batch_size)`` of each global batch of ``world_size * batch_size`` rows.
        The yielded tensors are reused by the next iteration when ``shuffle`` is on (staging buffers).
        """
        if self._ids is None:
            self._tokenise_all()
        n = len(self)
        gb = batch_size * world_size
        perm = torch.randperm(n, generator=generator) if shuffle else None
        if shuffle and self._stage is None or (self._stage is not None and self._stage[0].shape[0] != batch_size):
            self._stage = [torch.zeros((batch_size, self.max_length), dtype=torch.int32, pin_memory=self._pin)
                           for _ in range(3)]
        for g0 in range(0, n, gb):
            lo, hi = g0 + rank * batch_size, min(g0 + (rank + 1) * batch_size, n)
            if hi <= lo or (drop_last and (g0 + gb > n)):
                return
            if perm is None:
                yield tuple(t[lo:hi] for t in self._ids)
            else:
                idx = perm[lo:hi]
                out = []
                for t, st in zip(self._ids, self._stage):
                    torch.index_select(t, 0, idx, out=st[:hi - lo])
                    out.append(st[:hi - lo])
                yield tuple(out)
