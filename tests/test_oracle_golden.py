"""Pin the numpy oracle (oracle/two_tower_oracle.py) against vectors produced by RUNNING THE
REFERENCE (tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import two_tower_oracle as O

RTOL = 2e-5   # fp32 vs fp32, different summation order
ATOL = 2e-6


def close(a, b, rtol=RTOL, atol=ATOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = max(1.0, float(np.abs(b).max())) if b.size else 1.0
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol * scale)


def params_of(g, prefix="param_"):
    return {k[len(prefix):]: g[k] for k in g.files if k.startswith(prefix)}


TOWERS = [("mean_small", "mean"), ("mean_char", "mean"), ("mean_odd", "mean"),
          ("avg_proj", "avg"), ("avg_noproj", "avg")]


@pytest.mark.parametrize("name,kind", TOWERS)
def test_gather_and_pool(golden_dir, name, kind):
    g = np.load(os.path.join(golden_dir, f"tower_{name}.npz"))
    p = params_of(g)
    np.testing.assert_array_equal(O.embed_gather(g["q_ids"], p["embedding"]), g["gather_q"])
    pooled, count = O.masked_mean_pool(g["q_ids"], p["embedding"])
    close(pooled, g["pooled_q"])
    assert count[2] == 0 and np.all(pooled[2] == 0)        # all-pad row pools to exactly 0


@pytest.mark.parametrize("name,kind", TOWERS)
@pytest.mark.parametrize("loss", ["in_batch", "triplet", "multiple_negatives"])
def test_tower_forward_backward(golden_dir, name, kind, loss):
    g = np.load(os.path.join(golden_dir, f"tower_{name}.npz"))
    p = params_of(g)
    fwd = O.mean_tower_fwd if kind == "mean" else O.avg_tower_fwd
    bwd = O.mean_tower_bwd if kind == "mean" else O.avg_tower_bwd
    yq, cq = fwd(g["q_ids"], p)
    yd, cd = fwd(g["d_ids"], p)
    yn, cn = fwd(g["n_ids"], p)
    pre = loss + "_"
    close(yq, g[pre + "q_out"]); close(yd, g[pre + "d_out"]); close(yn, g[pre + "n_out"])
    nrm = np.linalg.norm(yq, axis=1)          # unit rows (avg_noproj: an all-pad row stays exactly 0)
    np.testing.assert_allclose(nrm[nrm > 0], 1.0, rtol=1e-5)
    if loss == "in_batch":
        L, _ = O.in_batch_loss(yq, yd, 0.1)
        dq, dd = O.in_batch_loss_bwd(yq, yd, 0.1)
        dn = None
    elif loss == "triplet":
        L = O.triplet_loss(yq, yd, yn, 0.2)
        dq, dd, dn = O.triplet_loss_bwd(yq, yd, yn, 0.2)
    else:
        negs = np.stack([yn, np.roll(yd, 1, 0), np.roll(yd, 2, 0)], axis=1)
        L = O.multiple_negatives_loss(yq, yd, negs, 0.1)
        dq, dd, dnegs = O.multiple_negatives_loss_bwd(yq, yd, negs, 0.1)
        close(dnegs, g[pre + "dnegs"], rtol=1e-4)
        dn = dnegs[:, 0]
        dd = dd + np.roll(dnegs[:, 1], -1, 0) + np.roll(dnegs[:, 2], -2, 0)
    close(L, g[pre + "loss"], rtol=1e-5)
    close(dq, g[pre + "dq_out"], rtol=1e-4); close(dd, g[pre + "dd_out"], rtol=1e-4)
    grads = bwd(dq, cq)
    for k, v in bwd(dd, cd).items():
        grads[k] = grads[k] + v
    if dn is not None:
        close(dn, g[pre + "dn_out"], rtol=1e-4)
        for k, v in bwd(dn, cn).items():
            grads[k] = grads[k] + v
    for k in p:
        close(grads[k], g[pre + "grad_" + k], rtol=2e-4, atol=1e-6)
    assert np.all(grads["embedding"][0] == 0)               # padding row gets zero grad


def test_losses_on_raw_rows(golden_dir):
    g = np.load(os.path.join(golden_dir, "losses_raw.npz"))
    q, p, n, negs = g["q"], g["p"], g["n"], g["negs"]
    close(O.in_batch_loss(q, p, 0.5)[0], g["in_batch_loss"])
    dq, dp = O.in_batch_loss_bwd(q, p, 0.5)
    close(dq, g["in_batch_dq"], rtol=1e-4); close(dp, g["in_batch_dp"], rtol=1e-4)
    close(O.triplet_loss(q, p, n, 0.3), g["triplet_loss"])
    dq, dp, dn = O.triplet_loss_bwd(q, p, n, 0.3)
    close(dq, g["triplet_dq"], rtol=1e-4); close(dp, g["triplet_dp"], rtol=1e-4)
    close(dn, g["triplet_dn"], rtol=1e-4)
    close(O.multiple_negatives_loss(q, p, negs, 0.2), g["multiple_negatives_loss"])
    dq, dp, dnegs = O.multiple_negatives_loss_bwd(q, p, negs, 0.2)
    close(dq, g["multiple_negatives_dq"], rtol=1e-4); close(dp, g["multiple_negatives_dp"], rtol=1e-4)
    close(dnegs, g["multiple_negatives_dnegs"], rtol=1e-4)


def test_in_batch_label_offset_is_block_of_global(golden_dir):
    """The multi-GPU generalisation: rank r's local loss with label_offset == rows r of the global loss."""
    g = np.load(os.path.join(golden_dir, "losses_raw.npz"))
    q, d = g["q"].astype(np.float64), g["p"].astype(np.float64)
    full, _ = O.in_batch_loss(q, d, 0.5)
    B = q.shape[0] // 2
    parts = [O.in_batch_loss(q[r * B:(r + 1) * B], d, 0.5, label_offset=r * B)[0] for r in range(2)]
    np.testing.assert_allclose(np.mean(parts), full, rtol=1e-12)


def test_search_topk(golden_dir):
    g = np.load(os.path.join(golden_dir, "search_topk.npz"))
    scores = O.search_scores(g["Q"], g["D"])
    close(scores, g["scores"], rtol=1e-5, atol=1e-6)
    vals, ids = O.topk_lower_index(g["scores"], 100)
    np.testing.assert_array_equal(ids, g["stable_indices"])
    np.testing.assert_array_equal(vals, g["stable_values"])
    # planted ties: rows 5,17,400,2999 identical -> lower index first
    assert list(ids[0][:4]) == [5, 17, 400, 2999]
    # torch.topk agrees up to tie order
    assert O.topk_ids_match(g["topk_indices"], g["topk_values"], ids, vals, g["scores"])
    np.testing.assert_array_equal(g["topk_values"], vals)


def test_search_small_end_to_end(golden_dir):
    """TwoTowerSearch.index_documents/search on strings, restated with the oracle."""
    with open(os.path.join(golden_dir, "search_small.json")) as f:
        gold = json.load(f)
    w = np.load(os.path.join(golden_dir, "search_small.npz"))
    vocab = gold["vocab"]

    def enc(text, L=64):
        ids = [vocab.get(c, 0) for c in text][:L]
        return ids + [0] * (L - len(ids))

    def tower(prefix, ids):
        p = dict(embedding=w[f"{prefix}__embedding__embedding__weight"],
                 w1=w[f"{prefix}__feed_forward__0__weight"], b1=w[f"{prefix}__feed_forward__0__bias"],
                 w2=w[f"{prefix}__feed_forward__2__weight"], b2=w[f"{prefix}__feed_forward__2__bias"])
        return O.mean_tower_fwd(np.array(ids), p)[0]

    D = tower("document_tower", [enc(d) for d in gold["docs"]])
    close(D, w["doc_embeddings"])
    for key, res in gold["results"].items():
        q, k = key.rsplit("|", 1)
        qv = tower("query_tower", [enc(q)])
        scores = O.search_scores(qv, D)
        vals, ids = O.topk_lower_index(scores, int(k))
        assert len(res) == min(int(k), len(gold["docs"]))
        close(vals[0], [r[1] for r in res], rtol=1e-5, atol=1e-6)
        # documents identical except across exact ties (docs 0 and 3 are the same string)
        assert [gold["docs"][i] for i in ids[0]] == [r[0] for r in res]


def test_adamw_and_train_steps(golden_dir):
    g = np.load(os.path.join(golden_dir, "train_triplet_3steps.npz"))
    names = ["embedding", "w1", "b1", "w2", "b2"]
    p = {k: g[f"init_{k}"] for k in names}
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v = {k: np.zeros_like(x) for k, x in p.items()}
    for step in range(3):
        outs, caches = [], []
        for nm in ["q_ids", "d_ids", "n_ids"]:
            y, c = O.mean_tower_fwd(g[f"step{step}_{nm}"], p)
            outs.append(y); caches.append(c)
        close(O.triplet_loss(*outs, 0.2), g[f"step{step}_loss"], rtol=1e-5)
        douts = O.triplet_loss_bwd(*outs, 0.2)
        grads = {k: 0 for k in names}
        for dy, c in zip(douts, caches):
            for k, val in O.mean_tower_bwd(dy, c).items():
                if k in grads:
                    grads[k] = grads[k] + val
        for k in names:
            p[k], m[k], v[k] = O.adamw_step(p[k], grads[k].astype(np.float32), m[k], v[k], step + 1)
            close(p[k], g[f"step{step}_{k}"], rtol=1e-5, atol=1e-6)


def test_torch_port_matches_reference_training(golden_dir):
    """oracle/torch_port.py (the CPU baseline that bench.py times) reproduces the reference's 3 steps."""
    import torch
    from oracle import torch_port as P
    g = np.load(os.path.join(golden_dir, "train_triplet_3steps.npz"))
    V, E = g["init_embedding"].shape
    H = g["init_w1"].shape[0]
    model = P.PortTwoTower(V, E, H, tied=True)
    model.query_tower.load_state_dict({
        "embedding.weight": torch.tensor(g["init_embedding"]),
        "feed_forward.0.weight": torch.tensor(g["init_w1"]), "feed_forward.0.bias": torch.tensor(g["init_b1"]),
        "feed_forward.2.weight": torch.tensor(g["init_w2"]), "feed_forward.2.bias": torch.tensor(g["init_b2"])})
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    for step in range(3):
        loss, _, _ = P.train_step(model, opt, "triplet", torch.tensor(g[f"step{step}_q_ids"]),
                                  torch.tensor(g[f"step{step}_d_ids"]), torch.tensor(g[f"step{step}_n_ids"]))
        close(loss, g[f"step{step}_loss"], rtol=1e-6)
        close(model.query_tower.feed_forward[2].weight.detach().numpy(), g[f"step{step}_w2"], rtol=1e-6)


def test_torch_port_search_matches_reference(golden_dir):
    import torch
    from oracle import torch_port as P
    g = np.load(os.path.join(golden_dir, "search_topk.npz"))
    v, i = P.search(torch.tensor(g["Q"][1:2]), torch.tensor(g["D"]), 100)
    np.testing.assert_array_equal(v.numpy(), g["topk_values"][1])
    np.testing.assert_array_equal(i.numpy(), g["topk_indices"][1])
