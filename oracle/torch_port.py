"""CPU port of the reference's PyTorch-eager hot path -- TEST / BASELINE INFRASTRUCTURE ONLY.

Used by ``bench.py``'s ``cpu_baseline`` leg and ``--impl reference`` arm (and nowhere in the
product).  The reference is Python and cannot travel to the GPU box (``/root/reference`` does
not exist there), so this file restates, with the same eager ATen ops in the same order, what
the reference executes per training step / per query, so that its timing on the box's host
cores is representative ("kind": "port"):

  * towers     -- twotower/encoders.py:62-77 (mask, embedding*mask, sum/(count+1e-9), MLP, normalize)
                  over nn.Embedding(V,E,padding_idx=0) twotower/embeddings.py:30
  * losses     -- twotower/losses.py:28-35 (triplet), :107-116 (in-batch)
  * train step -- twotower/train.py:120-154 (forward, loss, zero_grad/backward/step with
                  torch.optim.AdamW(lr) :359, monitoring cosines + 3x .item())
  * search     -- inference/search/two_tower.py:98-105 (cosine_similarity broadcast + topk)

Checked against the golden vectors in tests/test_oracle_golden.py::test_torch_port_*.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class PortTower(nn.Module):
    def __init__(self, embedding: nn.Embedding, hidden_dim: int):
        super().__init__()
        self.embedding = embedding
        e = embedding.embedding_dim
        self.feed_forward = nn.Sequential(nn.Linear(e, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, hidden_dim))

    def forward(self, input_ids):
        mask = (input_ids > 0).float().unsqueeze(-1)
        embeddings = self.embedding(input_ids) * mask
        pooled = embeddings.sum(1) / (mask.sum(1) + 1e-9)
        return F.normalize(self.feed_forward(pooled), dim=-1)


class PortTwoTower(nn.Module):
    def __init__(self, vocab_size: int, embedding_dim: int, hidden_dim: int, tied: bool = True):
        super().__init__()
        emb = nn.Embedding(vocab_size, embedding_dim, padding_idx=0)
        self.query_tower = PortTower(emb, hidden_dim)
        self.document_tower = self.query_tower if tied else PortTower(emb, hidden_dim)

    def forward(self, q, d=None, n=None):
        out = [self.query_tower(q)]
        if d is not None:
            out.append(self.document_tower(d))
        if n is not None:
            out.append(self.document_tower(n))
        return tuple(out)


def triplet_loss(q, p, n, margin=0.2):
    return F.relu(margin - F.cosine_similarity(q, p, dim=1) + F.cosine_similarity(q, n, dim=1)).mean()


def in_batch_loss(q, d, temperature=0.1):
    logits = torch.matmul(q, d.transpose(0, 1)) / temperature
    return F.cross_entropy(logits, torch.arange(q.shape[0], device=q.device))


def train_step(model, opt, loss_name, q_ids, d_ids, n_ids=None, temperature=0.1, margin=0.2):
    """One reference-shaped step incl. the monitoring host syncs (train.py:120-154)."""
    if loss_name == "in_batch":
        qv, dv = model(q_ids, d_ids)
        loss = in_batch_loss(qv, dv, temperature)
        nv = None
    else:
        qv, dv, nv = model(q_ids, d_ids, n_ids)
        loss = triplet_loss(qv, dv, nv, margin)
    opt.zero_grad()
    loss.backward()
    opt.step()
    with torch.no_grad():
        pos = F.cosine_similarity(qv, dv).mean().item()
        neg = F.cosine_similarity(qv, nv).mean().item() if nv is not None else 0.0
    return loss.item(), pos, neg


def search(q_emb, doc_embeddings, top_k):
    """two_tower.py:98-105 exactly: broadcast cosine + topk."""
    scores = F.cosine_similarity(q_emb.unsqueeze(1), doc_embeddings.unsqueeze(0), dim=2).squeeze(0)
    return torch.topk(scores, min(top_k, doc_embeddings.shape[0]))
