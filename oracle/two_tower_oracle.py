"""CPU oracle for the two-tower hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``two_towers_b200``) never imports anything under ``oracle/``.

What it is: a plain-numpy restatement (forward AND closed-form backward) of the
reference's PyTorch-eager algorithm for the path BASELINE.json names.  The arithmetic of
the reference lives in a third-party dependency that is not vendored under
``/root/reference``: **PyTorch ATen** (``requirements.txt:2`` pins ``torch>=2.2,<3.0``; the
container runs torch 2.11.0+cu128).  Each function below cites the reference call site
(file:line, relative to ``/root/reference``) whose behaviour it restates, plus the ATen
semantics that matter (epsilons, reductions).

Pinning: the reference's own tests hold no vectors for this path, so the
oracle is pinned against outputs of the reference itself executed in the build container:
``tests/golden/make_golden.py`` imports the reference modules from ``/root/reference``,
runs them (forward + autograd backward) on seeded inputs and commits the results under
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every oracle function against
those vectors.

All functions take / return numpy arrays.  ``dtype`` selects the working precision
(float32 mirrors the reference; float64 gives a tighter ground truth for tolerance checks).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "embed_gather", "masked_mean_pool", "masked_mean_pool_bwd",
    "linear", "normalize", "normalize_bwd", "layer_norm", "layer_norm_bwd",
    "mean_tower_fwd", "mean_tower_bwd", "avg_tower_fwd", "avg_tower_bwd",
    "cosine_similarity", "cosine_similarity_bwd",
    "in_batch_loss", "in_batch_loss_bwd", "triplet_loss", "triplet_loss_bwd",
    "multiple_negatives_loss", "multiple_negatives_loss_bwd",
    "search_scores", "topk_lower_index", "topk_ids_match", "adamw_step",
]


# --------------------------------------------------------------------------------------
# embeddings + pooling
# --------------------------------------------------------------------------------------
def embed_gather(ids: np.ndarray, table: np.ndarray) -> np.ndarray:
    """``LookupEmbedding.forward`` -> ``nn.Embedding`` row gather.

    Reference: twotower/embeddings.py:33-40 (``self.embedding(input_ids)``), table built at
    :30 with ``padding_idx=0`` (row 0 is zero-initialised and never receives gradient).
    """
    return table[ids]


def masked_mean_pool(ids: np.ndarray, table: np.ndarray, dtype=np.float32):
    """Gather + mask + mean over the sequence axis.

    Reference: twotower/encoders.py:62 (``mask = (input_ids > 0).float()``), :67
    (``embedding(ids) * mask``), :72 (``sum(1) / (mask.sum(1) + 1e-9)``); the same three lines
    at :125-138 for the avg_pool tower.  Note the mask is ``> 0`` (not ``!= padding_idx``) and
    the denominator keeps the ``+1e-9`` so that an all-pad row pools to exactly 0.

    Returns (pooled [B,E], count [B]).
    """
    ids = np.asarray(ids)
    mask = (ids > 0).astype(dtype)                       # [B,L]
    emb = table.astype(dtype)[ids] * mask[..., None]     # [B,L,E]
    count = mask.sum(1)                                  # [B]
    pooled = emb.sum(1) / (count[:, None] + dtype(1e-9))
    return pooled.astype(dtype), count.astype(dtype)


def masked_mean_pool_bwd(ids: np.ndarray, d_pooled: np.ndarray, vocab_size: int,
                         dtype=np.float32) -> np.ndarray:
    """Dense embedding gradient of :func:`masked_mean_pool`.

    Reference: autograd of twotower/encoders.py:67-72 ending in ATen
    ``embedding_dense_backward`` (triggered by ``loss.backward()`` twotower/train.py:138):
    ``dW[v] = sum_{(b,l): ids[b,l]==v, v>0} d_pooled[b] / (count_b + 1e-9)``; row
    ``padding_idx=0`` receives exactly zero.
    """
    ids = np.asarray(ids)
    B, L = ids.shape
    mask = (ids > 0)
    count = mask.sum(1).astype(dtype)
    g = d_pooled.astype(dtype) / (count[:, None] + dtype(1e-9))   # [B,E]
    dW = np.zeros((vocab_size, d_pooled.shape[1]), dtype=dtype)
    rows = np.repeat(np.arange(B), L)[mask.reshape(-1)]
    np.add.at(dW, ids.reshape(-1)[mask.reshape(-1)], g[rows])
    dW[0] = 0
    return dW


# --------------------------------------------------------------------------------------
# dense pieces
# --------------------------------------------------------------------------------------
def linear(x, weight, bias):
    """``nn.Linear``: ``x @ weight.T + bias`` (weight is [out,in]).  twotower/encoders.py:38-42."""
    return x @ weight.T + bias


def normalize(z, eps=1e-12):
    """``F.normalize(z, dim=-1)``: ``z / max(||z||_2, eps)``.  twotower/encoders.py:77,150."""
    n = np.sqrt((z * z).sum(-1, keepdims=True))
    return z / np.maximum(n, z.dtype.type(eps))


def normalize_bwd(dy, z, eps=1e-12):
    """Backward of :func:`normalize`: ``dz = (dy - y (y.dy)) / max(||z||, eps)``.

    (ATen clamps the norm, so below ``eps`` the gradient is simply ``dy / eps``.)
    """
    n = np.sqrt((z * z).sum(-1, keepdims=True))
    nc = np.maximum(n, z.dtype.type(eps))
    y = z / nc
    inner = (dy * y).sum(-1, keepdims=True)
    return np.where(n > eps, (dy - y * inner) / nc, dy / nc)


def layer_norm(x, gamma, beta, eps=1e-5):
    """``nn.LayerNorm(H)`` (biased variance, eps 1e-5).  twotower/encoders.py:103."""
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + x.dtype.type(eps))
    xhat = (x - mu) * rstd
    return xhat * gamma + beta, xhat, rstd


def layer_norm_bwd(dy, xhat, rstd, gamma):
    H = dy.shape[-1]
    dgamma = (dy * xhat).sum(0)
    dbeta = dy.sum(0)
    dxhat = dy * gamma
    dx = (dxhat - dxhat.mean(-1, keepdims=True)
          - xhat * (dxhat * xhat).mean(-1, keepdims=True)) * rstd
    return dx, dgamma, dbeta


# --------------------------------------------------------------------------------------
# towers
# --------------------------------------------------------------------------------------
def mean_tower_fwd(ids, params, dtype=np.float32):
    """``MeanPoolingTower.forward`` -- twotower/encoders.py:50-81.

    params: dict with ``embedding`` [V,E], ``w1`` [H,E], ``b1`` [H], ``w2`` [H,H], ``b2`` [H]
    (= ``embedding.embedding.weight``, ``feed_forward.0.{weight,bias}``,
    ``feed_forward.2.{weight,bias}``).  Returns (y [B,H] unit rows, cache for bwd).
    """
    p = {k: np.asarray(v, dtype=dtype) for k, v in params.items()}
    pooled, count = masked_mean_pool(ids, p["embedding"], dtype)
    a1 = linear(pooled, p["w1"], p["b1"])          # :39 Linear(E,H)
    h1 = np.maximum(a1, 0)                          # :40 ReLU
    z = linear(h1, p["w2"], p["b2"])                # :41 Linear(H,H)
    y = normalize(z)                                # :77
    return y, dict(ids=np.asarray(ids), pooled=pooled, count=count, a1=a1, h1=h1, z=z, p=p)


def mean_tower_bwd(dy, cache, dtype=np.float32):
    """Closed-form backward of :func:`mean_tower_fwd` (what autograd does for the reference)."""
    p = cache["p"]
    dz = normalize_bwd(dy.astype(dtype), cache["z"])
    dw2 = dz.T @ cache["h1"]
    db2 = dz.sum(0)
    dh1 = dz @ p["w2"]
    da1 = dh1 * (cache["a1"] > 0)
    dw1 = da1.T @ cache["pooled"]
    db1 = da1.sum(0)
    dpooled = da1 @ p["w1"]
    demb = masked_mean_pool_bwd(cache["ids"], dpooled, p["embedding"].shape[0], dtype)
    return dict(embedding=demb, w1=dw1, b1=db1, w2=dw2, b2=db2, pooled=dpooled)


def avg_tower_fwd(ids, params, dtype=np.float32):
    """``AveragePoolingTower.forward`` in eval mode / dropout p=0 -- twotower/encoders.py:113-155.

    params: ``embedding`` [V,E]; when H != E also ``w`` [H,E], ``b`` [H] (``projection.0``),
    ``gamma`` [H], ``beta`` [H] (``projection.2`` LayerNorm).  Dropout (:102) is the identity
    in eval mode; train-mode dropout draws from torch's RNG and is not restated.
    """
    p = {k: np.asarray(v, dtype=dtype) for k, v in params.items()}
    pooled, count = masked_mean_pool(ids, p["embedding"], dtype)
    cache = dict(ids=np.asarray(ids), pooled=pooled, count=count, p=p)
    if "w" in p:
        a = linear(pooled, p["w"], p["b"])
        ln, xhat, rstd = layer_norm(a, p["gamma"], p["beta"])
        cache.update(a=a, xhat=xhat, rstd=rstd, z=ln)
        z = ln
    else:
        z = pooled
        cache.update(z=z)
    return normalize(z), cache


def avg_tower_bwd(dy, cache, dtype=np.float32):
    p = cache["p"]
    dz = normalize_bwd(dy.astype(dtype), cache["z"])
    out = {}
    if "w" in p:
        da, dgamma, dbeta = layer_norm_bwd(dz, cache["xhat"], cache["rstd"], p["gamma"])
        out.update(gamma=dgamma, beta=dbeta, w=da.T @ cache["pooled"], b=da.sum(0))
        dpooled = da @ p["w"]
    else:
        dpooled = dz
    out["pooled"] = dpooled
    out["embedding"] = masked_mean_pool_bwd(cache["ids"], dpooled, p["embedding"].shape[0], dtype)
    return out


# --------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------
def cosine_similarity(x, y, eps=1e-8):
    """``F.cosine_similarity(x, y, dim=-1)``: ``x.y / (max(||x||,eps) * max(||y||,eps))``.

    Used at twotower/losses.py:28-29,74, inference/search/two_tower.py:98-102,
    twotower/train.py:145-150.
    """
    e = x.dtype.type(eps)
    nx = np.maximum(np.sqrt((x * x).sum(-1)), e)
    ny = np.maximum(np.sqrt((y * y).sum(-1)), e)
    return (x * y).sum(-1) / (nx * ny)


def cosine_similarity_bwd(dc, x, y, eps=1e-8):
    """Gradients of cos wrt x and y (norms above eps): ``y/(nx ny) - c x/nx^2``."""
    e = x.dtype.type(eps)
    nx = np.maximum(np.sqrt((x * x).sum(-1, keepdims=True)), e)
    ny = np.maximum(np.sqrt((y * y).sum(-1, keepdims=True)), e)
    c = (x * y).sum(-1, keepdims=True) / (nx * ny)
    dc = dc[..., None]
    dx = dc * (y / (nx * ny) - c * x / (nx * nx))
    dy = dc * (x / (nx * ny) - c * y / (ny * ny))
    return dx, dy


def _log_softmax(logits):
    m = logits.max(-1, keepdims=True)
    lse = m + np.log(np.exp(logits - m).sum(-1, keepdims=True))
    return logits - lse, lse[..., 0]


def in_batch_loss(q, d, temperature=0.1, label_offset=0):
    """``in_batch_sampled_softmax_loss`` -- twotower/losses.py:88-118.

    ``S = q @ d.T`` (:107); ``logits = S / temperature`` (:110); labels = arange(B) (:113);
    ``F.cross_entropy`` mean over rows (:116).  ``label_offset`` is the multi-GPU
    generalisation (row i's positive is column ``i + label_offset``); 0 == the reference.
    Returns (loss, lse [Bq]).
    """
    t = q.dtype.type(temperature)
    logits = (q @ d.T) / t
    logp, lse = _log_softmax(logits)
    idx = np.arange(q.shape[0])
    return -logp[idx, idx + label_offset].mean(), lse


def in_batch_loss_bwd(q, d, temperature=0.1, label_offset=0, grad=1.0):
    """dL/dq, dL/dd for :func:`in_batch_loss`: ``G = (softmax(logits) - onehot) / (B t)``."""
    t = q.dtype.type(temperature)
    logits = (q @ d.T) / t
    logp, _ = _log_softmax(logits)
    G = np.exp(logp)
    idx = np.arange(q.shape[0])
    G[idx, idx + label_offset] -= 1
    G *= q.dtype.type(grad) / (q.shape[0] * t)
    return G @ d, G.T @ q


def triplet_loss(q, p, n, margin=0.2):
    """``contrastive_triplet_loss`` -- twotower/losses.py:9-44: ``relu(m - cos(q,p) + cos(q,n)).mean()``."""
    sp, sn = cosine_similarity(q, p), cosine_similarity(q, n)
    return np.maximum(q.dtype.type(margin) - sp + sn, 0).mean()


def triplet_loss_bwd(q, p, n, margin=0.2, grad=1.0):
    sp, sn = cosine_similarity(q, p), cosine_similarity(q, n)
    active = ((q.dtype.type(margin) - sp + sn) > 0).astype(q.dtype) * q.dtype.type(grad / q.shape[0])
    dq1, dp = cosine_similarity_bwd(-active, q, p)
    dq2, dn = cosine_similarity_bwd(active, q, n)
    return dq1 + dq2, dp, dn


def multiple_negatives_loss(q, p, negs, temperature=0.1):
    """``multiple_negatives_loss`` -- twotower/losses.py:47-85: cos(q,[p;negs]) / t, CE vs label 0."""
    docs = np.concatenate([p[:, None, :], negs], axis=1)                 # :71  [B,N+1,H]
    sims = cosine_similarity(np.broadcast_to(q[:, None, :], docs.shape), docs)   # :68,:74
    logp, _ = _log_softmax(sims / q.dtype.type(temperature))              # :77
    return -logp[:, 0].mean()                                             # :80-83


def multiple_negatives_loss_bwd(q, p, negs, temperature=0.1, grad=1.0):
    docs = np.concatenate([p[:, None, :], negs], axis=1)
    qe = np.broadcast_to(q[:, None, :], docs.shape)
    sims = cosine_similarity(qe, docs)
    logp, _ = _log_softmax(sims / q.dtype.type(temperature))
    G = np.exp(logp)
    G[:, 0] -= 1
    G *= q.dtype.type(grad) / (q.shape[0] * q.dtype.type(temperature))    # d loss / d sims
    dqe, ddocs = cosine_similarity_bwd(G, qe, docs)
    return dqe.sum(1), ddocs[:, 0], ddocs[:, 1:]


# --------------------------------------------------------------------------------------
# retrieval
# --------------------------------------------------------------------------------------
def search_scores(q, D):
    """``TwoTowerSearch.search`` scoring -- inference/search/two_tower.py:98-102:
    ``F.cosine_similarity(q[1,1,H], D[1,N,H], dim=2)`` -> [N] (q may be [nq,H] -> [nq,N])."""
    q2 = np.atleast_2d(q)
    return cosine_similarity(q2[:, None, :], D[None, :, :])


def topk_lower_index(scores, k):
    """Top-k with the BASELINE tie rule: descending score, ties -> LOWER index first.

    Reference: ``torch.topk(similarity_scores, min(top_k, N))`` at
    inference/search/two_tower.py:105.  torch.topk's CPU tie order is arbitrary (SURVEY.md
    section 2c), so the tie oracle is a stable descending sort.
    """
    scores = np.atleast_2d(scores)
    k = min(k, scores.shape[1])
    order = np.argsort(-scores, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(scores, order, 1), order.astype(np.int64)


def topk_ids_match(ids, scores, ref_ids, ref_scores, all_scores, rtol=1e-5, atol=1e-6):
    """True iff ``ids`` equals ``ref_ids`` except at positions whose score ties (within
    tolerance) with a neighbour / the k-th boundary -- the check BASELINE.json asks for."""
    ids, ref_ids = np.asarray(ids), np.asarray(ref_ids)
    if ids.shape != ref_ids.shape:
        return False
    tol = atol + rtol * np.abs(ref_scores)
    if not np.all(np.abs(np.asarray(scores) - ref_scores) <= tol + 1e-6):
        return False
    bad = ids != ref_ids
    if not bad.any():
        return True
    # a mismatching id is acceptable only if its true score equals the reference score there
    rows = np.nonzero(bad)[0]
    true_scores = all_scores[rows, ids[bad]]
    return bool(np.all(np.abs(true_scores - ref_scores[bad]) <= tol[bad]))


# --------------------------------------------------------------------------------------
# optimizer
# --------------------------------------------------------------------------------------
def adamw_step(p, g, m, v, step, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01):
    """One ``torch.optim.AdamW`` step (defaults; twotower/train.py:359 ``AdamW(params, lr)``).

    ATen ``_single_tensor_adamw``: decoupled decay ``p *= 1 - lr*wd``; moments; bias
    corrections ``1 - beta^step``; ``p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)``.
    ``step`` is the 1-based step number.  Returns new (p, m, v).
    """
    p = p * (1 - lr * weight_decay)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = np.sqrt(v) / np.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p.astype(g.dtype), m.astype(g.dtype), v.astype(g.dtype)
