"""Ranking evaluation behind the reference's ``evaluate_model`` interface.

Reference: twotower/evaluate.py -- ``mean_reciprocal_rank`` :16, ``precision_at_k`` :39, ``recall_at_k`` :65,
``ndcg_at_k`` :95 (scikit-learn ``ndcg_score`` on the ranked relevance list), ``evaluate_model`` :126-236 (encode the
query and its candidate documents with the towers, rank by cosine similarity, average the metrics over queries).

Same signatures, same metric definitions and result keys.  What changes is where the work happens: queries and
candidate lists are tokenised in batches (``encode_batch``), embedded by the B200 towers in large batches, and ranked on
the device by the same fused scan + exact top-k kernel the search path uses (ties -> lower index); only the
[n_candidates] ranking of each query returns to the host, where the (tiny) metric arithmetic stays numpy.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple, Union

import numpy as np
import torch

from . import ops

Number = Union[int, float]


def mean_reciprocal_rank(relevance_scores: Sequence[Number]) -> float:
    """1 / rank of the first relevant (== 1) entry of a ranked relevance list; 0 if none (evaluate.py:16-37)."""
    rel = np.asarray(relevance_scores)
    hits = np.flatnonzero(rel == 1)
    return 0.0 if hits.size == 0 else 1.0 / float(hits[0] + 1)


def precision_at_k(relevance_scores: Sequence[Number], k: int) -> float:
    """Mean relevance of the top k, short lists padded with zeros (evaluate.py:39-63)."""
    rel = np.asarray(relevance_scores, dtype=np.float64)
    top = np.zeros(k, dtype=np.float64)
    n = min(k, rel.shape[0])
    top[:n] = rel[:n]
    return float(top.mean())


def recall_at_k(relevance_scores: Sequence[Number], k: int, total_relevant: Number) -> float:
    """Relevant entries in the top k over all relevant ones; 0 when there are none (evaluate.py:65-93)."""
    if total_relevant == 0:
        return 0.0
    rel = np.asarray(relevance_scores)
    return float(rel[:k].sum() / total_relevant)


def ndcg_at_k(relevance_scores: Sequence[Number], k: int) -> float:
    """The reference's formulation (evaluate.py:95-124): scikit-learn's ndcg_score with y_true = the relevance list
    sorted descending and y_score = the list in ranked order, both zero-padded to k."""
    from sklearn.metrics import ndcg_score
    rel = np.asarray(relevance_scores)
    ideal = np.sort(rel)[::-1]
    if rel.shape[0] < k:
        pad = k - rel.shape[0]
        ideal, rel = np.pad(ideal, (0, pad)), np.pad(rel, (0, pad))
    return float(ndcg_score(ideal.reshape(1, -1), rel.reshape(1, -1), k=k))


@torch.no_grad()
def rank_candidates(model, tokenizer, query: str, documents: Sequence[str], max_len: int = 64, batch_size: int = 8192,
                    device="cuda") -> np.ndarray:
    """Indices of `documents` by descending cosine similarity to `query` (ties: lower index first)."""
    dev = torch.device(device)
    q_ids = tokenizer.encode_batch([query], max_len).to(dev)
    q = model.query_tower(q_ids)
    chunks = []
    for i in range(0, len(documents), batch_size):
        ids = tokenizer.encode_batch(list(documents[i:i + batch_size]), max_len).to(dev)
        chunks.append(model.document_tower(ids))
    d = torch.cat(chunks, 0) if len(chunks) > 1 else chunks[0]
    n = d.shape[0]
    if n <= ops.TT_TOPK_MAX:
        _, order = ops.topk_scan(d.contiguous(), q.contiguous(), n, cosine=True)       # exact full ranking on the device
        return order[0].cpu().numpy()
    # candidate lists longer than the kernel's k limit: exact full ranking of every block of TT_TOPK_MAX candidates on the
    # device (scores + ids from the scan kernel), then one host merge of the sorted blocks by (score desc, index asc) --
    # no eager-PyTorch scoring path
    d = d.contiguous()
    ss, ii = [], []
    for lo in range(0, n, ops.TT_TOPK_MAX):
        hi = min(n, lo + ops.TT_TOPK_MAX)
        s_blk, i_blk = ops.topk_scan(d[lo:hi], q.contiguous(), hi - lo, cosine=True, id_offset=lo)
        ss.append(s_blk[0]); ii.append(i_blk[0])
    s_all, i_all = torch.cat(ss).cpu().numpy(), torch.cat(ii).cpu().numpy()
    return i_all[np.lexsort((i_all, -s_all.astype(np.float64)))]


def evaluate_model(model, test_data: List[Tuple[str, List[str], List[int]]], tokenizer,
                   metrics: List[str] = ["precision", "recall", "mrr", "ndcg"], k_values: List[int] = [1, 5, 10],
                   batch_size: int = 32, device: str = "cuda") -> Dict[str, float]:
    """evaluate.py:126-236.  `batch_size` is accepted for signature compatibility; documents are embedded in batches of
    max(batch_size, 8192)."""
    model.eval()
    model = model.to(device)
    per_p, per_r, per_mrr, per_n = [], [], [], []
    for query, documents, relevance in test_data:
        order = rank_candidates(model, tokenizer, query, documents, batch_size=max(batch_size, 8192), device=device)
        ranked = np.asarray(relevance)[order]
        total = np.sum(relevance)
        per_p.append([precision_at_k(ranked, k) for k in k_values])
        per_r.append([recall_at_k(ranked, k, total) for k in k_values])
        per_mrr.append(mean_reciprocal_rank(ranked))
        per_n.append([ndcg_at_k(ranked, k) for k in k_values])
    out: Dict[str, float] = {}
    if "precision" in metrics:
        for i, k in enumerate(k_values):
            out[f"precision@{k}"] = float(np.mean([p[i] for p in per_p]))
    if "recall" in metrics:
        for i, k in enumerate(k_values):
            out[f"recall@{k}"] = float(np.mean([r[i] for r in per_r]))
    if "mrr" in metrics:
        out["mrr"] = float(np.mean(per_mrr))
    if "ndcg" in metrics:
        for i, k in enumerate(k_values):
            out[f"ndcg@{k}"] = float(np.mean([n[i] for n in per_n]))
    return out
