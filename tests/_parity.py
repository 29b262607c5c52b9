"""Shared parity helpers for the GPU tests: the stated tolerances, the error metrics, and the fp64 oracle of one
whole training step (towers -> loss -> every parameter gradient).

Tolerances (BASELINE.json north_star): fp32 mode rel 1e-5, bf16 mode rel 2e-2.  Metrics:
  max-abs  : max|a - b| / max|b|          (per tensor)
  norm-wise: ||a - b||_2 / ||b||_2        (per tensor)
``check`` asserts BOTH at the given rtol and returns / prints the measured values so the GPU log carries the
per-tensor error next to every assertion.

ReLU gates: a pre-activation within bf16 rounding of zero can take the other branch in bf16 mode.  The gate is a
discontinuity of the reference function itself (an eps perturbation of the input moves the gradient by O(sqrt(eps))
in norm), so step-level bf16 comparisons evaluate the oracle WITH THE KERNEL'S OWN GATES (read from its saved h1) and
separately bound the fraction of gates that differ from the fp64 ones.
"""
from __future__ import annotations

import numpy as np
import torch

from oracle import two_tower_oracle as O

FP32_RTOL = 1e-5
BF16_RTOL = 2e-2


def to_np(a):
    if torch.is_tensor(a):
        return a.detach().float().cpu().numpy().astype(np.float64)
    return np.asarray(a, np.float64)


def errors(a, b):
    a, b = to_np(a), to_np(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.isfinite(a).all(), "non-finite output"
    d = a - b
    return (float(np.abs(d).max()) / max(float(np.abs(b).max()), 1e-30),
            float(np.linalg.norm(d)) / max(float(np.linalg.norm(b)), 1e-30))


def check(a, b, rtol, what="", log=None):
    mx, fro = errors(a, b)
    line = f"    parity {what:<28s} max-abs {mx:.2e}  norm-wise {fro:.2e}  (tol {rtol:.0e})"
    print(line)
    if log is not None:
        log.append(line)
    assert mx <= rtol and fro <= rtol, f"{what}: max-abs {mx:.3e} / norm-wise {fro:.3e} exceed {rtol:.0e}"
    return mx, fro


def tower_params(tower):
    """numpy copies keyed like the oracle's mean_tower_fwd."""
    g = lambda t: t.detach().float().cpu().numpy().astype(np.float64)
    return dict(embedding=g(tower.embedding.embedding.weight), w1=g(tower.feed_forward[0].weight),
                b1=g(tower.feed_forward[0].bias), w2=g(tower.feed_forward[2].weight), b2=g(tower.feed_forward[2].bias))


def tower_grads(tower):
    g = lambda t: t.grad.detach().float().cpu().numpy().astype(np.float64)
    return dict(embedding=g(tower.embedding.embedding.weight), w1=g(tower.feed_forward[0].weight),
                b1=g(tower.feed_forward[0].bias), w2=g(tower.feed_forward[2].weight), b2=g(tower.feed_forward[2].bias))


def bf16_round(a):
    return torch.tensor(np.asarray(a), dtype=torch.float32).bfloat16().double().numpy()


def oracle_step(pq, pd, q_ids, d_ids, n_ids=None, loss="in_batch", temperature=0.1, margin=0.2, gates=None,
                groups=None, quantize_y=False):
    """fp64 oracle of one step (reference semantics: twotower/train.py:120-139 on the given batch).

    pq / pd: parameter dicts of the query / document tower (pd is pq when tied; the embedding is always shared).
    gates: optional (gq, gd, gn) boolean [B,H] arrays replacing the oracle's own ReLU gates (see module docstring).
    groups: optional list of (lo, hi) row ranges -- in-batch negatives are then LOCAL to each range and the loss is the
            mean over ranges (data-parallel 'local negatives' / DDP semantics); None = one global batch.
    quantize_y: round the tower outputs to bf16 before the loss (what TT_PREC_BF16 feeds the loss kernels by definition):
            separates the kernels' own arithmetic error from the input quantisation of the mode when the loss gradient is
            ill-conditioned (nearly collinear outputs: dq_i = (sum_j P_ij d_j - d_i) / (tau B) is a small difference).
    Returns (loss, grads_q, grads_d, flip_fraction): gradients per tower (for tied towers grads_d is grads_q = the sum).
    """
    f = np.float64
    tied = pd is pq
    ids = [np.asarray(q_ids), np.asarray(d_ids)] + ([np.asarray(n_ids)] if n_ids is not None else [])
    ps = [pq, pd, pd][:len(ids)]
    ys, caches, flips = [], [], []
    for k, (i, p) in enumerate(zip(ids, ps)):
        y, c = O.mean_tower_fwd(i, p, f)
        if gates is not None and gates[k] is not None:
            own = c["a1"] > 0
            g = np.asarray(gates[k], bool)
            flips.append(float((own != g).mean()))
            c["a1"] = np.where(g, 1.0, -1.0)
        ys.append(bf16_round(y) if quantize_y else y); caches.append(c)
    B = ids[0].shape[0]
    if loss == "in_batch":
        rng = groups if groups is not None else [(0, B)]
        total, dq, dd = 0.0, np.zeros_like(ys[0]), np.zeros_like(ys[1])
        for lo, hi in rng:
            l, _ = O.in_batch_loss(ys[0][lo:hi], ys[1][lo:hi], temperature)
            a, b = O.in_batch_loss_bwd(ys[0][lo:hi], ys[1][lo:hi], temperature)
            total += l / len(rng); dq[lo:hi] = a / len(rng); dd[lo:hi] = b / len(rng)
        dys = [dq, dd]
    elif loss == "triplet":
        total = O.triplet_loss(ys[0], ys[1], ys[2], margin)
        dys = list(O.triplet_loss_bwd(ys[0], ys[1], ys[2], margin))
    else:
        raise ValueError(loss)
    gs = [O.mean_tower_bwd(dy, c, f) for dy, c in zip(dys, caches)]
    keys = ("embedding", "w1", "b1", "w2", "b2")
    gq = {k: gs[0][k].copy() for k in keys}
    gd = {k: sum(g[k] for g in gs[1:]) for k in keys}
    emb = gq["embedding"] + gd["embedding"]                      # one shared table (encoders.py:265,270)
    gq["embedding"] = gd["embedding"] = emb
    if tied:
        for k in keys[1:]:
            gq[k] = gq[k] + gd[k]
        gd = gq
    return float(total), gq, gd, (max(flips) if flips else 0.0)


def trainer_gates(tr):
    """The ReLU gates the trainer's last step actually used, as (gq, gd[, gn]) boolean arrays (bf16 mode only)."""
    B, P = tr.B, tr.passes
    rows = []
    for gi, (_, r0, nr) in enumerate(tr.groups):
        h = tr.h1_bf16[gi]
        rows.append((h.float() > 0).cpu().numpy())
    allr = np.concatenate(rows, 0)
    return tuple(allr[k * B:(k + 1) * B] for k in range(P))
