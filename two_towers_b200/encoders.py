"""Towers behind the reference's TOWER_REGISTRY (twotower/encoders.py).

Same constructors, submodule / parameter names (``feed_forward.0/2``, ``projection.0/2``) and
outputs (unit-norm [B,H]) as the reference, so checkpoints are interchangeable; the math is
two kernels per tower pass: fused gather+masked-mean-pool (K1) and the tower MLP /
projection+LayerNorm with the L2 normalise folded in (K3 / K3').
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .embeddings import BaseEmbedding


class BaseTower(nn.Module):
    """encoders.py:12-22"""

    def __init__(self, embedding: BaseEmbedding, hidden_dim: int):
        super().__init__()
        self.embedding = embedding
        self.hidden_dim = hidden_dim
        self.precision = None          # None -> ops default; 'fp32' | 'bf16'

    def log_params(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)


class MeanPoolingTower(BaseTower):
    """encoders.py:25-81: mask -> gather*mask -> sum/len -> Linear-ReLU-Linear -> F.normalize."""

    def __init__(self, embedding: BaseEmbedding, hidden_dim: int):
        super().__init__(embedding, hidden_dim)
        e = embedding.embedding_dim
        self.feed_forward = nn.Sequential(nn.Linear(e, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, hidden_dim))

    def forward(self, input_ids: torch.Tensor) -> torch.Tensor:
        pooled = self.embedding.pooled(input_ids)                                  # encoders.py:62-72
        l1, l2 = self.feed_forward[0], self.feed_forward[2]
        if torch.is_grad_enabled() and (pooled.requires_grad or l1.weight.requires_grad):
            return ops.MlpFn.apply(pooled, l1.weight, l1.bias, l2.weight, l2.bias, self.precision)
        return ops.mlp_fwd(pooled, l1.weight.detach(), l1.bias.detach(), l2.weight.detach(), l2.bias.detach(),
                           self.precision)[0]                                      # encoders.py:77


class AveragePoolingTower(BaseTower):
    """encoders.py:84-155: same pool; if H != E: Linear -> Dropout -> LayerNorm; F.normalize."""
    _instances = 0

    def __init__(self, embedding: BaseEmbedding, hidden_dim: int, dropout: float = 0.1):
        super().__init__(embedding, hidden_dim)
        e = embedding.embedding_dim
        self.has_projection = hidden_dim != e
        self.dropout_p = float(dropout)
        # every tower instance (and every rank) draws its own dropout stream: untied query / document towers must not
        # share masks (nn.Dropout modules never do)
        AveragePoolingTower._instances += 1
        self._seed = (0x5EED + 0x9E3779B97F4A7C15 * AveragePoolingTower._instances) % (1 << 64)
        if self.has_projection:
            self.projection = nn.Sequential(nn.Linear(e, hidden_dim), nn.Dropout(dropout), nn.LayerNorm(hidden_dim))

    def forward(self, input_ids: torch.Tensor) -> torch.Tensor:
        pooled = self.embedding.pooled(input_ids)                                  # encoders.py:125-138
        if self.has_projection:
            lin, ln = self.projection[0], self.projection[2]
            args = (lin.weight, lin.bias, ln.weight, ln.bias)
        else:
            args = (None, None, None, None)
        drop_on = self.training and self.has_projection and self.dropout_p > 0
        if drop_on:
            self._seed = (self._seed * 6364136223846793005 + 1442695040888963407) % (1 << 64)
        if torch.is_grad_enabled() and (pooled.requires_grad or
                                        (self.has_projection and args[0].requires_grad)):
            return ops.ProjLnFn.apply(pooled, *args, self.has_projection, self.dropout_p, drop_on, self._seed, self.precision)
        args = tuple(None if a is None else a.detach() for a in args)
        return ops.proj_ln_fwd(pooled, *args, self.has_projection, self.dropout_p, drop_on, self._seed, self.precision)[0]


class TwoTower(nn.Module):
    """encoders.py:158-224"""

    def __init__(self, query_tower: BaseTower, document_tower: BaseTower = None, tied_weights: bool = False):
        super().__init__()
        self.query_tower = query_tower
        if tied_weights:
            self.document_tower = query_tower
        else:
            self.document_tower = document_tower if document_tower is not None else query_tower

    def forward(self, query_input, document_input=None, negative_input=None):
        q = self.query_tower(query_input)
        d = self.document_tower(document_input) if document_input is not None else None
        n = self.document_tower(negative_input) if negative_input is not None else None
        if n is not None:
            return q, d, n
        if d is not None:
            return q, d
        return q

    def encode_query(self, query_input):
        return self.query_tower(query_input)

    def encode_document(self, document_input):
        return self.document_tower(document_input)


TOWER_REGISTRY = {"mean": MeanPoolingTower, "avg_pool": AveragePoolingTower}


def build_tower(name: str, embedding: BaseEmbedding, **kwargs) -> BaseTower:
    """encoders.py:234-248"""
    if name not in TOWER_REGISTRY:
        raise ValueError(f"Unknown tower architecture: {name}. Available options: {list(TOWER_REGISTRY.keys())}")
    return TOWER_REGISTRY[name](embedding=embedding, **kwargs)


def build_two_tower(tower_name: str, embedding: BaseEmbedding, hidden_dim: int, tied_weights: bool = False,
                    **kwargs) -> TwoTower:
    """encoders.py:251-271 (the embedding object is shared by both towers even when untied)."""
    query_tower = build_tower(tower_name, embedding, hidden_dim=hidden_dim, **kwargs)
    document_tower = None if tied_weights else build_tower(tower_name, embedding, hidden_dim=hidden_dim, **kwargs)
    return TwoTower(query_tower, document_tower, tied_weights=tied_weights)
