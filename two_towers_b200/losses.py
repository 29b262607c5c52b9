"""Loss functions behind the reference's LOSS_REGISTRY (twotower/losses.py).

Identical names, signatures and defaults (margin 0.2, temperature 0.1); each is one fused
forward kernel + one fused backward kernel.  ``in_batch_sampled_softmax_loss`` additionally
accepts the reference train loop's 3-positional call ``loss_fn(q, pos, neg)`` (train.py:133) --
the third positional is ignored when it is a tensor (the reference itself raises a TypeError
there) -- and the multi-GPU ``label_offset`` / ``loss_scale`` generalisation.
"""
from __future__ import annotations

from functools import partial
from typing import Callable

import torch

from . import ops


def contrastive_triplet_loss(q_emb, d_pos_emb, d_neg_emb, margin: float = 0.2) -> torch.Tensor:
    """losses.py:9-44"""
    return ops.TripletLossFn.apply(q_emb, d_pos_emb, d_neg_emb, float(margin))


def multiple_negatives_loss(q_emb, d_pos_emb, d_neg_embs, temperature: float = 0.1) -> torch.Tensor:
    """losses.py:47-85"""
    return ops.MultiNegLossFn.apply(q_emb, d_pos_emb, d_neg_embs, float(temperature))


def in_batch_sampled_softmax_loss(q_emb, d_emb, temperature=0.1, *, label_offset: int = 0, loss_scale=None,
                                  precision=None) -> torch.Tensor:
    """losses.py:88-118"""
    if torch.is_tensor(temperature):           # called as loss_fn(q, pos, neg): neg is unused by this loss
        temperature = 0.1
    return ops.InBatchLossFn.apply(q_emb, d_emb, float(temperature), int(label_offset), loss_scale, precision)


def _in_batch_adapter(q_emb, d_emb, d_neg_emb=None, *, temperature: float = 0.1, **kw):
    return in_batch_sampled_softmax_loss(q_emb, d_emb, temperature, **kw)


LOSS_REGISTRY = {
    "triplet": contrastive_triplet_loss,
    "multiple_negatives": multiple_negatives_loss,
    "in_batch": in_batch_sampled_softmax_loss,
}


def build(name: str, **kwargs) -> Callable:
    """losses.py:129-150.  ``in_batch`` with kwargs returns the (q, pos, neg)-tolerant adapter."""
    if name not in LOSS_REGISTRY:
        raise ValueError(f"Unknown loss function: {name}. Available options: {list(LOSS_REGISTRY.keys())}")
    if name == "in_batch":
        return partial(_in_batch_adapter, **kwargs) if kwargs else in_batch_sampled_softmax_loss
    fn = LOSS_REGISTRY[name]
    return partial(fn, **kwargs) if kwargs else fn
