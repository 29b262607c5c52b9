"""Embedding layers behind the reference's registry (twotower/embeddings.py).

``LookupEmbedding`` keeps the reference contract -- ctor ``(vocab_size, embedding_dim,
padding_idx=0)``, parameter ``embedding.weight`` (so reference checkpoints load with
``load_state_dict``), ``forward(ids[B,L]) -> [B,L,E]`` (embeddings.py:33-40) -- and adds
``pooled(ids) -> [B,E]``: the fused gather + masked-mean-pool kernel the towers call, so the
[B,L,E] tensor of encoders.py:67 never exists.
"""
from __future__ import annotations

from abc import ABC

import torch
import torch.nn as nn

from . import ops


class BaseEmbedding(nn.Module, ABC):
    """embeddings.py:10-16"""

    def __init__(self, vocab_size: int, embedding_dim: int, padding_idx: int = 0):
        super().__init__()
        self.vocab_size = vocab_size
        self.embedding_dim = embedding_dim
        self.padding_idx = padding_idx

    def log_params(self):
        return self.vocab_size * self.embedding_dim

    # fused path used by the towers
    def pooled(self, input_ids: torch.Tensor) -> torch.Tensor:
        w = self.embedding.weight
        if w.requires_grad and torch.is_grad_enabled():
            return ops.EmbedPoolFn.apply(input_ids, w)
        return ops.embed_pool_fwd(input_ids, w.detach())[0]

    def forward(self, input_ids: torch.Tensor) -> torch.Tensor:
        w = self.embedding.weight
        if w.requires_grad and torch.is_grad_enabled():
            return ops.EmbedGatherFn.apply(input_ids, w)
        return ops.embed_gather(input_ids, w.detach())


class LookupEmbedding(BaseEmbedding):
    """nn.Embedding(V, E, padding_idx=0) storage (N(0,1) init, row 0 zero), B200 kernels for the math."""

    def __init__(self, vocab_size: int, embedding_dim: int, padding_idx: int = 0):
        super().__init__(vocab_size, embedding_dim, padding_idx)
        self.embedding = nn.Embedding(vocab_size, embedding_dim, padding_idx=padding_idx)


class PretrainedEmbedding(BaseEmbedding):
    """Table initialised from a [V,E] array (row 0 forced to zero), frozen unless trainable."""

    def __init__(self, vectors, trainable: bool = False, padding_idx: int = 0):
        vectors = torch.as_tensor(vectors, dtype=torch.float32)
        super().__init__(vectors.shape[0], vectors.shape[1], padding_idx)
        self.embedding = nn.Embedding.from_pretrained(vectors.clone(), freeze=not trainable, padding_idx=padding_idx)


class FrozenWord2Vec(PretrainedEmbedding):
    """embeddings.py:43-84: gensim KeyedVectors with a zero PAD row prepended, frozen."""

    def __init__(self, kv_path: str, vocab_size: int = None, embedding_dim: int = None, padding_idx: int = 0):
        try:
            import gensim
            import numpy as np
        except ImportError:
            raise ImportError("Please install gensim to use FrozenWord2Vec embedding: pip install gensim")
        kv = gensim.models.KeyedVectors.load(kv_path, mmap="r")
        vecs = torch.cat([torch.zeros(1, kv.vector_size), torch.tensor(np.array(kv.vectors), dtype=torch.float)])
        super().__init__(vecs, trainable=False, padding_idx=padding_idx)
        self.kv = kv


class GloVeEmbedding(BaseEmbedding):
    """embeddings.py:87-155: rows 1..min(len(glove),V)-1 initialised from gensim-data vectors."""

    def __init__(self, vocab_size: int, embedding_dim: int = None, model_name: str = "glove-wiki-gigaword-50",
                 trainable: bool = False, padding_idx: int = 0):
        try:
            import gensim.downloader as api
        except ImportError:
            raise ImportError("Please install gensim to use GloVeEmbedding: pip install gensim")
        model = api.load(model_name)
        embedding_dim = model.vector_size if embedding_dim is None else embedding_dim
        super().__init__(vocab_size, embedding_dim, padding_idx)
        self.embedding = nn.Embedding(vocab_size, embedding_dim, padding_idx=padding_idx)
        weight = torch.zeros_like(self.embedding.weight)
        for i in range(1, min(len(model.index_to_key), vocab_size)):
            weight[i] = torch.tensor(model[model.index_to_key[i - 1]], dtype=weight.dtype)
        with torch.no_grad():
            self.embedding.weight.copy_(weight)
        self.embedding.weight.requires_grad = trainable
        self.model = model


REGISTRY = {"lookup": LookupEmbedding, "word2vec": FrozenWord2Vec, "glove": GloVeEmbedding}


def build(name: str, vocab_size: int, **kwargs) -> BaseEmbedding:
    """embeddings.py:166-180"""
    if name not in REGISTRY:
        raise ValueError(f"Unknown embedding: {name}. Available options: {list(REGISTRY.keys())}")
    return REGISTRY[name](vocab_size=vocab_size, **kwargs)
