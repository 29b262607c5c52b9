"""GPU parity tests (run with ``-m gpu`` on a B200): the CUDA path, called through the C ABI
(``two_towers_b200.ops`` -> ctypes -> libtt_b200.so), against the oracle and the golden
vectors produced by running the reference.

Tolerances (BASELINE.json north_star): fp32 mode rel 1e-5 for pooled embeddings / loss /
gradients (checked as |a-b| <= 1e-5 * max|b| + tiny atol, since individual gradient entries
cross zero); bf16 mode rel 2e-2; top-k ids identical except at score ties (lower index wins).
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import two_tower_oracle as O

pytestmark = pytest.mark.gpu

FP32_RTOL = 1e-5
DEV = "cuda"


def cu(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


def close(a, b, rtol=FP32_RTOL, atol_scale=2e-6):
    a = a.detach().float().cpu().numpy().astype(np.float64) if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = max(float(np.abs(b).max()), 1e-30) if b.size else 1.0
    err = float(np.abs(a - b).max()) if b.size else 0.0
    assert err <= rtol * scale + atol_scale * scale, f"max err {err:.3e} vs scale {scale:.3e} (rel {err / scale:.2e})"


def params_of(g, prefix="param_"):
    return {k[len(prefix):]: g[k] for k in g.files if k.startswith(prefix)}


def make_ids(rng, B, L, V, zipf=False):
    ids = np.zeros((B, L), np.int64)
    for b in range(B):
        n = rng.integers(0 if b % 17 == 3 else 1, L + 1)
        if zipf:
            x = np.minimum(rng.zipf(1.07, n), V - 1)
        else:
            x = rng.integers(1, V, n)
        ids[b, :n] = x
    if B > 1 and L > 2:
        ids[1, 1] = 0
    return ids


def test_native_library_is_loaded_and_counts_launches():
    import two_towers_b200 as tt
    before = tt._lib.launch_count()
    tt.ops.embed_pool_fwd(cu(np.ones((2, 3), np.int64)), cu(np.ones((4, 4), np.float32)))
    assert tt._lib.launch_count() > before
    maps = open("/proc/self/maps").read()
    assert "libtt_b200.so" in maps
    assert tt._lib.load().tt_require_sm100(0) == 0


# ------------------------------------------------------------------------------------------
# K1 / K2
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("V,E,B,L", [(128, 64, 64, 64), (5000, 300, 33, 32), (57, 20, 5, 7), (40, 7, 9, 5),
                                     (1000, 512, 16, 16), (50, 4, 3, 1), (3000, 1024, 4, 9)])
@pytest.mark.parametrize("idt", [torch.int64, torch.int32])
def test_embed_pool_fwd_and_gather(V, E, B, L, idt):
    import two_towers_b200 as tt
    rng = np.random.default_rng(V + E)
    table = rng.standard_normal((V, E)).astype(np.float32)
    table[0] = 0
    ids = make_ids(rng, B, L, V)
    pooled, inv_len, pb = tt.ops.embed_pool_fwd(cu(ids, idt), cu(table), want_bf16=True)
    ref, cnt = O.masked_mean_pool(ids, table)
    close(pooled, ref)
    close(inv_len, 1.0 / (cnt + np.float32(1e-9)), rtol=1e-6)
    close(pb, ref, rtol=1e-2)
    empty = cnt == 0
    if empty.any():
        assert torch.all(pooled[cu(empty)] == 0)                    # all-pad row pools to exactly 0
    out = tt.ops.embed_gather(cu(ids, idt), cu(table))
    np.testing.assert_array_equal(out.cpu().numpy(), O.embed_gather(ids, table))


def test_embed_pool_golden(golden_dir):
    import two_towers_b200 as tt
    for name in ("mean_small", "mean_char", "mean_odd", "avg_proj"):
        g = np.load(os.path.join(golden_dir, f"tower_{name}.npz"))
        pooled, _, _ = tt.ops.embed_pool_fwd(cu(g["q_ids"]), cu(g["param_embedding"]))
        close(pooled, g["pooled_q"])
        np.testing.assert_array_equal(tt.ops.embed_gather(cu(g["q_ids"]), cu(g["param_embedding"])).cpu().numpy(),
                                      g["gather_q"])


@pytest.mark.parametrize("V,E,B,L,zipf", [(128, 64, 256, 64, False), (34, 64, 32, 64, False), (1024, 20, 50, 9, False),
                                          (5000, 300, 64, 32, True), (400_000, 300, 512, 32, True),
                                          (2000, 7, 40, 5, False), (1500, 64, 700, 64, True)])
def test_embed_pool_bwd_deterministic(V, E, B, L, zipf):
    import two_towers_b200 as tt
    rng = np.random.default_rng(V * 31 + E)
    ids = make_ids(rng, B, L, V, zipf)
    g = rng.standard_normal((B, E)).astype(np.float32)
    cnt = (ids > 0).sum(1).astype(np.float32)
    inv_len = (1.0 / (cnt + np.float32(1e-9))).astype(np.float32)
    d1 = tt.ops.embed_pool_bwd(cu(ids), cu(inv_len), cu(g), V)
    d2 = tt.ops.embed_pool_bwd(cu(ids), cu(inv_len), cu(g), V)
    assert torch.equal(d1, d2)                                      # bitwise reproducible
    ref = O.masked_mean_pool_bwd(ids, g.astype(np.float64), V, np.float64)
    close(d1, ref)
    assert torch.all(d1[0] == 0)                                    # padding row
    untouched = np.setdiff1d(np.arange(V), np.unique(ids))
    if untouched.size:
        assert torch.all(d1[cu(untouched)] == 0)


# ------------------------------------------------------------------------------------------
# K3 / K3' towers against the golden vectors of the reference
# ------------------------------------------------------------------------------------------
def build_model(tt, g, kind, tied=True, precision=None):
    p = params_of(g)
    V, E = p["embedding"].shape
    emb = tt.embeddings.build("lookup", V, embedding_dim=E)
    if kind == "mean":
        H = p["w1"].shape[0]
        model = tt.build_two_tower("mean", emb, hidden_dim=H, tied_weights=tied)
    else:
        H = p["w"].shape[0] if "w" in p else E
        model = tt.build_two_tower("avg_pool", emb, hidden_dim=H, tied_weights=tied, dropout=0.0)
    sd = {"embedding.embedding.weight": p["embedding"]}
    if kind == "mean":
        sd.update({"feed_forward.0.weight": p["w1"], "feed_forward.0.bias": p["b1"],
                   "feed_forward.2.weight": p["w2"], "feed_forward.2.bias": p["b2"]})
    elif "w" in p:
        sd.update({"projection.0.weight": p["w"], "projection.0.bias": p["b"],
                   "projection.2.weight": p["gamma"], "projection.2.bias": p["beta"]})
    model.query_tower.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})     # reference key names
    model.query_tower.precision = precision
    return model.to(DEV), p


TOWERS = [("mean_small", "mean"), ("mean_char", "mean"), ("mean_odd", "mean"), ("avg_proj", "avg"),
          ("avg_noproj", "avg")]
GRAD_KEYS = {"embedding": "embedding.embedding.weight", "w1": "feed_forward.0.weight", "b1": "feed_forward.0.bias",
             "w2": "feed_forward.2.weight", "b2": "feed_forward.2.bias", "w": "projection.0.weight",
             "b": "projection.0.bias", "gamma": "projection.2.weight", "beta": "projection.2.bias"}


@pytest.mark.parametrize("name,kind", TOWERS)
@pytest.mark.parametrize("loss", ["in_batch", "triplet", "multiple_negatives"])
def test_tower_loss_backward_matches_reference(golden_dir, name, kind, loss):
    """Full module path: registries -> towers -> loss -> loss.backward(), vs reference autograd."""
    import two_towers_b200 as tt
    g = np.load(os.path.join(golden_dir, f"tower_{name}.npz"))
    model, p = build_model(tt, g, kind)
    model.eval()
    q_ids, d_ids, n_ids = cu(g["q_ids"]), cu(g["d_ids"]), cu(g["n_ids"])
    qv, dv, nv = model(q_ids, d_ids, n_ids)
    for t in (qv, dv, nv):
        t.retain_grad()
    pre = loss + "_"
    close(qv, g[pre + "q_out"]); close(dv, g[pre + "d_out"]); close(nv, g[pre + "n_out"])
    if loss == "in_batch":
        L = tt.losses.build("in_batch", temperature=0.1)(qv, dv, nv)       # reference loop call shape
    elif loss == "triplet":
        L = tt.losses.build("triplet", margin=0.2)(qv, dv, nv)
    else:
        negs = torch.stack([nv, dv.roll(1, 0), dv.roll(2, 0)], dim=1)
        negs.retain_grad()
        L = tt.losses.build("multiple_negatives", temperature=0.1)(qv, dv, negs)
    L.backward()
    close(L, g[pre + "loss"])
    close(qv.grad, g[pre + "dq_out"], rtol=2e-5); close(dv.grad, g[pre + "dd_out"], rtol=2e-5)
    if loss == "multiple_negatives":
        close(negs.grad, g[pre + "dnegs"], rtol=2e-5)
    named = dict(model.query_tower.named_parameters())
    for k in p:
        close(named[GRAD_KEYS[k]].grad, g[pre + "grad_" + k], rtol=5e-5)
    assert torch.all(named["embedding.embedding.weight"].grad[0] == 0)


def test_losses_on_raw_rows_match_reference(golden_dir):
    import two_towers_b200 as tt
    g = np.load(os.path.join(golden_dir, "losses_raw.npz"))
    q, p, n, negs = (cu(g[k]).requires_grad_() for k in ("q", "p", "n", "negs"))
    for nm, fn in [("in_batch", lambda: tt.in_batch_sampled_softmax_loss(q, p, temperature=0.5)),
                   ("triplet", lambda: tt.contrastive_triplet_loss(q, p, n, margin=0.3)),
                   ("multiple_negatives", lambda: tt.multiple_negatives_loss(q, p, negs, temperature=0.2))]:
        for t in (q, p, n, negs):
            t.grad = None
        L = fn()
        L.backward()
        close(L, g[f"{nm}_loss"])
        for tn, t in (("q", q), ("p", p), ("n", n), ("negs", negs)):
            if f"{nm}_d{tn}" in g.files:
                close(t.grad, g[f"{nm}_d{tn}"], rtol=2e-5)


@pytest.mark.parametrize("Bq,Bd,H,off,temp", [(5, 5, 24, 0, 0.1), (64, 64, 256, 0, 0.1), (100, 257, 64, 57, 0.05),
                                              (257, 300, 130, 3, 1.0), (1024, 1024, 256, 0, 0.1), (96, 768, 256, 96 * 3, 0.1)])
def test_inbatch_ce_fp32_vs_oracle(Bq, Bd, H, off, temp):
    import two_towers_b200 as tt
    rng = np.random.default_rng(Bq + Bd + H)
    q = O.normalize(rng.standard_normal((Bq, H))).astype(np.float32)
    d = O.normalize(rng.standard_normal((Bd, H))).astype(np.float32)
    loss, lse, pm = tt.ops.inbatch_ce_fwd(cu(q), cu(d), temp, off, precision="fp32", want_pos_mean=True)
    q64, d64 = q.astype(np.float64), d.astype(np.float64)
    rl, rlse = O.in_batch_loss(q64, d64, temp, off)
    close(loss, rl); close(lse, rlse)
    close(pm, np.mean([(q64[i] * d64[i + off]).sum() for i in range(Bq)]), rtol=1e-4)
    gout = cu(np.float32(0.5))
    dq, dd = tt.ops.inbatch_ce_bwd(cu(q), cu(d), lse, temp, off, grad_out=gout, precision="fp32")
    rdq, rdd = O.in_batch_loss_bwd(q64, d64, temp, off, grad=0.5)
    close(dq, rdq, rtol=2e-5); close(dd, rdd, rtol=2e-5)
    dq2, dd2 = tt.ops.inbatch_ce_bwd(cu(q), cu(d), lse, temp, off, grad_out=gout, precision="fp32")
    assert torch.equal(dq, dq2) and torch.equal(dd, dd2)                      # deterministic


def test_inbatch_ce_full_size_known_answers():
    """B=4096, d=256 (the BASELINE shape) through size-independent properties."""
    import two_towers_b200 as tt
    B, H = 4096, 256
    rng = np.random.default_rng(1)
    q = torch.nn.functional.normalize(torch.randn(B, H, device=DEV), dim=-1)
    v = torch.nn.functional.normalize(torch.randn(1, H, device=DEV), dim=-1)
    d_same = v.expand(B, H).contiguous()
    # all documents identical -> uniform softmax: loss == log(B) exactly, dq == 0
    loss, lse, _ = tt.ops.inbatch_ce_fwd(q, d_same, 0.1, precision="fp32")
    assert abs(loss.item() - np.log(B)) < 1e-4
    dq, dd = tt.ops.inbatch_ce_bwd(q, d_same, lse, 0.1, precision="fp32")
    assert dq.abs().max().item() < 1e-7
    # block decomposition: mean of label_offset block losses == full loss (the multi-GPU identity)
    d = torch.nn.functional.normalize(torch.randn(B, H, device=DEV), dim=-1)
    full, lse_full, _ = tt.ops.inbatch_ce_fwd(q, d, 0.1, precision="fp32")
    parts = [tt.ops.inbatch_ce_fwd(q[r * 1024:(r + 1) * 1024].contiguous(), d, 0.1, label_offset=r * 1024,
                                   precision="fp32")[0].item() for r in range(4)]
    assert abs(np.mean(parts) - full.item()) < 2e-5 * abs(full.item())
    # gradient rows sum rule: sum_j G_ij = 0  =>  sum_i dd_i . w == -(dq . ...) checked via linearity:
    dq, dd = tt.ops.inbatch_ce_bwd(q, d, lse_full, 0.1, precision="fp32")
    # d loss along a common shift of all logits is zero: <dq, q> + <dd, d> both equal sum G_ij S_ij / ...; they match
    assert abs((dq * q).sum().item() - (dd * d).sum().item()) < 1e-4
    # compare a 256-row slab against the float64 oracle
    rl, _ = O.in_batch_loss(q[:256].double().cpu().numpy(), d.double().cpu().numpy(), 0.1)
    part0 = tt.ops.inbatch_ce_fwd(q[:256].contiguous(), d, 0.1, precision="fp32")[0].item()
    assert abs(part0 - rl) < 1e-5 * abs(rl)


# ------------------------------------------------------------------------------------------
# K7 search
# ------------------------------------------------------------------------------------------
def test_topk_golden_with_planted_ties(golden_dir):
    import two_towers_b200 as tt
    g = np.load(os.path.join(golden_dir, "search_topk.npz"))
    for cosine in (True, False):
        s, i = tt.ops.topk_scan(cu(g["D"]), cu(g["Q"]), 100, cosine=cosine)
        assert O.topk_ids_match(i.cpu().numpy(), s.cpu().numpy(), g["stable_indices"], g["stable_values"], g["scores"])
        assert list(i[0, :4].cpu().numpy()) == [5, 17, 400, 2999]         # exact ties -> lower index first
        sv = s.cpu().numpy()
        assert np.all(sv[:, :-1] >= sv[:, 1:])


@pytest.mark.parametrize("N,H,nq,k,dtype", [(200_000, 256, 3, 100, "fp32"), (200_000, 256, 2, 100, "bf16"),
                                             (50_000, 64, 1, 1, "fp32"), (4097, 128, 5, 1024, "fp32"),
                                             (1000, 50, 2, 10, "fp32"), (130, 256, 1, 130, "fp32"),
                                             (3, 32, 1, 3, "fp32"), (70_000, 512, 2, 37, "bf16"),
                                             (9999, 72, 1, 5, "bf16")])
def test_topk_scan_vs_oracle(N, H, nq, k, dtype):
    import two_towers_b200 as tt
    rng = np.random.default_rng(N + H)
    D = O.normalize(rng.standard_normal((N, H))).astype(np.float32)
    Q = O.normalize(rng.standard_normal((nq, H))).astype(np.float32)
    Dt = cu(D)
    if dtype == "bf16":
        Dt = tt.ops.cast_bf16(Dt)
        D = Dt.float().cpu().numpy()                      # the oracle scores the same (rounded) index
    for cosine in (False, True):
        s, i = tt.ops.topk_scan(Dt, cu(Q), k, cosine=cosine, id_offset=1000)
        scores = (Q.astype(np.float64) @ D.astype(np.float64).T) if not cosine else \
            O.search_scores(Q.astype(np.float64), D.astype(np.float64))
        rv, ri = O.topk_lower_index(scores, k)
        assert O.topk_ids_match(i.cpu().numpy() - 1000, s.cpu().numpy(), ri, rv, scores, rtol=2e-6, atol=2e-6)


def test_topk_duplicates_everywhere_lower_index_wins():
    import two_towers_b200 as tt
    D = np.zeros((10_000, 64), np.float32); D[:, 0] = 1.0          # every row identical
    Q = np.zeros((1, 64), np.float32); Q[0, 0] = 1.0
    s, i = tt.ops.topk_scan(cu(D), cu(Q), 100, cosine=False)
    assert i[0].cpu().tolist() == list(range(100))
    assert torch.all(s == 1.0)


def test_topk_merge_matches_single_scan():
    import two_towers_b200 as tt
    rng = np.random.default_rng(5)
    N, H, k = 30_000, 128, 50
    D = O.normalize(rng.standard_normal((N, H))).astype(np.float32)
    D[20_000] = D[7]
    Q = O.normalize(rng.standard_normal((4, H))).astype(np.float32); Q[0] = D[7]
    full_s, full_i = tt.ops.topk_scan(cu(D), cu(Q), k, cosine=False)
    bounds = [0, 7000, 7001, 19_000, N]
    parts = [tt.ops.topk_scan(cu(D[a:b]), cu(Q), min(k, b - a), cosine=False, id_offset=a) for a, b in zip(bounds, bounds[1:])]
    S = torch.full((len(parts), 4, k), float("-inf"), device=DEV); I = torch.full((len(parts), 4, k), -1, dtype=torch.int64, device=DEV)
    for r, (s, i) in enumerate(parts):
        S[r, :, :s.shape[1]] = s; I[r, :, :i.shape[1]] = i
    ms, mi = tt.ops.topk_merge(S, I)
    assert torch.equal(mi, full_i) and torch.equal(ms, full_s)
    assert mi[0, :2].tolist() == [7, 20_000]


def test_search_end_to_end_matches_reference(golden_dir, tmp_path):
    """TwoTowerSearch on strings: same documents/scores as the reference class."""
    import two_towers_b200 as tt
    gold = json.load(open(os.path.join(golden_dir, "search_small.json")))
    w = np.load(os.path.join(golden_dir, "search_small.npz"))
    tok = tt.CharTokeniser(); tok.string_to_index = gold["vocab"]
    tok.index_to_string = {i: c for c, i in gold["vocab"].items()}
    emb = tt.embeddings.build("lookup", tok.vocab_size, embedding_dim=16)
    model = tt.build_two_tower("mean", emb, hidden_dim=32, tied_weights=False)
    model.load_state_dict({k.replace("__", "."): torch.tensor(w[k]) for k in w.files if k != "doc_embeddings"})
    for idx_dtype in ("fp32", "bf16"):
        s = tt.TwoTowerSearch(model, tok, device=DEV, index_dtype=idx_dtype)
        with pytest.raises(ValueError):
            s.search("x")
        s.index_documents(gold["docs"])
        assert s.document_embeddings.is_contiguous() and s.document_embeddings.shape == (len(gold["docs"]), 32)
        if idx_dtype == "fp32":
            close(s.document_embeddings, w["doc_embeddings"])
        for key, res in gold["results"].items():
            q, k = key.rsplit("|", 1)
            out = s.search(q, top_k=int(k))
            assert len(out) == len(res) and set(out[0]) == {"document", "score"}
            tol = 1e-5 if idx_dtype == "fp32" else 2e-2
            np.testing.assert_allclose([o["score"] for o in out], [r[1] for r in res], rtol=tol, atol=tol)
            if idx_dtype == "fp32":
                assert [o["document"] for o in out] == [r[0] for r in res]
    s = tt.TwoTowerSearch(model, tok, device=DEV)
    s.index_documents(gold["docs"])
    s.save_index(tmp_path / "idx.pkl")
    s2 = tt.TwoTowerSearch(model, tok, device=DEV)
    s2.load_index(tmp_path / "idx.pkl")
    assert s2.search("cat", 3) == s.search("cat", 3)
    import pickle
    blob = pickle.load(open(tmp_path / "idx.pkl", "rb"))
    assert set(blob) == {"embeddings", "documents"} and isinstance(blob["embeddings"], np.ndarray)
    # raw, mmap-able format: bit-exact round trip for both index dtypes, and a row-range (shard) load
    for idx_dtype in ("fp32", "bf16"):
        sa = tt.TwoTowerSearch(model, tok, device=DEV, index_dtype=idx_dtype)
        sa.index_documents(gold["docs"])
        sa.save_index_raw(str(tmp_path / f"raw_{idx_dtype}"))
        sb = tt.TwoTowerSearch(model, tok, device=DEV, index_dtype=idx_dtype)
        sb.load_index_raw(str(tmp_path / f"raw_{idx_dtype}"))
        assert torch.equal(sb.document_embeddings, sa.document_embeddings) and sb.documents == list(gold["docs"])
        assert sb.search("cat", 3) == sa.search("cat", 3)
        n = len(gold["docs"])
        sc = tt.TwoTowerSearch(model, tok, device=DEV, index_dtype=idx_dtype)
        sc.load_index_raw(str(tmp_path / f"raw_{idx_dtype}"), rows=range(1, n))
        assert torch.equal(sc.document_embeddings, sa.document_embeddings[1:]) and sc.row_offset == 1


def test_evaluate_model_matches_numpy_ranking(golden_dir):
    """evaluate_model (device ranking through the scan / top-k kernel) == the reference procedure restated in numpy on the
    same embeddings: cosine scores, descending sort, the metric helpers (pinned against the reference in the CPU suite)."""
    import two_towers_b200 as tt
    from two_towers_b200 import evaluate as E
    gold = json.load(open(os.path.join(golden_dir, "search_small.json")))
    w = np.load(os.path.join(golden_dir, "search_small.npz"))
    tok = tt.CharTokeniser(); tok.string_to_index = gold["vocab"]
    tok.index_to_string = {i: c for c, i in gold["vocab"].items()}
    emb = tt.embeddings.build("lookup", tok.vocab_size, embedding_dim=16)
    model = tt.build_two_tower("mean", emb, hidden_dim=32, tied_weights=False).to(DEV)
    model.load_state_dict({k.replace("__", "."): torch.tensor(w[k]) for k in w.files if k != "doc_embeddings"})
    docs = list(gold["docs"])
    rng = np.random.default_rng(5)
    test_data = []
    for qi, q in enumerate(["cat", "dog food", "quantum", "the"]):
        cand = [docs[j] for j in rng.permutation(len(docs))[:max(3, len(docs) - qi)]]
        rel = rng.integers(0, 2, len(cand)).tolist()
        test_data.append((q, cand, rel))
    got = E.evaluate_model(model, test_data, tok, k_values=[1, 3, 5], device=DEV)
    # numpy restatement
    P, R, M, N = [], [], [], []
    with torch.no_grad():
        for q, cand, rel in test_data:
            qe = model.query_tower(tok.encode_batch([q], 64).to(DEV)).double().cpu().numpy()[0]
            de = model.document_tower(tok.encode_batch(cand, 64).to(DEV)).double().cpu().numpy()
            sc = de @ qe / (np.linalg.norm(de, axis=1) * np.linalg.norm(qe) + 1e-30)
            order = np.argsort(-sc, kind="stable")
            ranked = np.asarray(rel)[order]
            P.append([E.precision_at_k(ranked, k) for k in (1, 3, 5)]); R.append([E.recall_at_k(ranked, k, sum(rel)) for k in (1, 3, 5)])
            M.append(E.mean_reciprocal_rank(ranked)); N.append([E.ndcg_at_k(ranked, k) for k in (1, 3, 5)])
    for i, k in enumerate((1, 3, 5)):
        assert abs(got[f"precision@{k}"] - np.mean([p[i] for p in P])) < 1e-12
        assert abs(got[f"recall@{k}"] - np.mean([r[i] for r in R])) < 1e-12
        assert abs(got[f"ndcg@{k}"] - np.mean([n[i] for n in N])) < 1e-12
    assert abs(got["mrr"] - np.mean(M)) < 1e-12
    assert set(got) == {"precision@1", "precision@3", "precision@5", "recall@1", "recall@3", "recall@5", "mrr", "ndcg@1", "ndcg@3", "ndcg@5"}


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_topk_full_size_planted_rows(dtype):
    """N = 10M x 256 (the BASELINE index): plant 100 rows that must win, at known positions."""
    import two_towers_b200 as tt
    N, H, k = 10_000_000, 256, 100
    gen = torch.Generator(device=DEV).manual_seed(7)
    D = torch.empty(N, H, device=DEV, dtype=torch.float32)
    for a in range(0, N, 1_000_000):
        D[a:a + 1_000_000] = torch.nn.functional.normalize(torch.randn(1_000_000, H, device=DEV, generator=gen), dim=-1)
    q = torch.nn.functional.normalize(torch.randn(1, H, device=DEV, generator=gen), dim=-1)
    pos = torch.randperm(N, device=DEV, generator=gen)[:k].sort().values
    pos[0], pos[-1] = 0, N - 1
    # planted rows: q scaled so their dot products are 2.0 - j*1e-3 (strictly ordered), far above random (~0.2)
    strengths = 2.0 - 1e-3 * torch.arange(k, device=DEV, dtype=torch.float32)
    perm = torch.randperm(k, device=DEV, generator=gen)
    D[pos] = strengths[perm, None] * q
    index = D if dtype == "fp32" else tt.ops.cast_bf16(D)
    s, i = tt.ops.topk_scan(index, q, k, cosine=False)
    expect = pos[torch.argsort(perm)]
    if dtype == "fp32":
        assert torch.equal(i[0], expect)
        assert torch.allclose(s[0], strengths, rtol=1e-5)
    else:
        assert set(i[0].tolist()) == set(pos.tolist())          # bf16 rounding may reorder 1e-3-spaced scores
    # sortedness + idempotence: searching the top-k rows alone gives the same list
    assert torch.all(s[0, :-1] >= s[0, 1:])
    sub = index[i[0]].contiguous()
    s2, i2 = tt.ops.topk_scan(sub, q, k, cosine=False)
    assert torch.equal(i[0][i2[0]], i[0]) and torch.allclose(s2, s)


# ------------------------------------------------------------------------------------------
# optimizer + fused trainer vs the reference's 3-step training fixtures
# ------------------------------------------------------------------------------------------
def test_adamw_matches_oracle():
    import two_towers_b200 as tt
    rng = np.random.default_rng(0)
    n = 10_007
    p = rng.standard_normal(n).astype(np.float32); m = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
    tp, tm, tv = cu(p), cu(m), cu(v)
    step = torch.zeros(2, dtype=torch.int64, device=DEV)
    for t in range(1, 6):
        g = rng.standard_normal(n).astype(np.float32)
        tt.ops.adamw_step(tp, cu(g), tm, tv, step)
        p, m, v = O.adamw_step(p.astype(np.float64), g.astype(np.float64), m.astype(np.float64), v.astype(np.float64), t)
        close(tp, p, rtol=5e-6); close(tm, m, rtol=5e-6); close(tv, v, rtol=5e-6)
        p, m, v = p.astype(np.float32), m.astype(np.float32), v.astype(np.float32)
    assert step.tolist() == [5, 0]


@pytest.mark.parametrize("graph", [False, True])
def test_fused_trainer_triplet_tied_matches_reference_3_steps(golden_dir, graph):
    import two_towers_b200 as tt
    g = np.load(os.path.join(golden_dir, "train_triplet_3steps.npz"))
    V, E = g["init_embedding"].shape; H = g["init_w1"].shape[0]
    emb = tt.embeddings.build("lookup", V, embedding_dim=E)
    model = tt.build_two_tower("mean", emb, hidden_dim=H, tied_weights=True)
    model.query_tower.load_state_dict({"embedding.embedding.weight": torch.tensor(g["init_embedding"]),
                                       "feed_forward.0.weight": torch.tensor(g["init_w1"]), "feed_forward.0.bias": torch.tensor(g["init_b1"]),
                                       "feed_forward.2.weight": torch.tensor(g["init_w2"]), "feed_forward.2.bias": torch.tensor(g["init_b2"])})
    model = model.to(DEV)
    B, L = g["step0_q_ids"].shape
    tr = tt.FusedTrainer(model, loss="triplet", margin=0.2, lr=1e-3, batch_size=B, max_len=L, precision="fp32",
                         use_cuda_graph=graph)
    for step in range(3):
        loss = tr.step(torch.tensor(g[f"step{step}_q_ids"]), torch.tensor(g[f"step{step}_d_ids"]),
                       torch.tensor(g[f"step{step}_n_ids"]))
        close(loss, g[f"step{step}_loss"], rtol=2e-5)
        sd = model.query_tower.state_dict()
        for k, key in (("embedding", "embedding.embedding.weight"), ("w1", "feed_forward.0.weight"),
                       ("b1", "feed_forward.0.bias"), ("w2", "feed_forward.2.weight"), ("b2", "feed_forward.2.bias")):
            close(sd[key], g[f"step{step}_{k}"], rtol=2e-5)


@pytest.mark.parametrize("graph", [False, True])
def test_fused_trainer_inbatch_untied_matches_reference_3_steps(golden_dir, graph):
    import two_towers_b200 as tt
    g = np.load(os.path.join(golden_dir, "train_inbatch_untied_3steps.npz"))
    V, E = g["init_q_embedding"].shape; H = g["init_q_w1"].shape[0]
    emb = tt.embeddings.build("lookup", V, embedding_dim=E)
    model = tt.build_two_tower("mean", emb, hidden_dim=H, tied_weights=False)
    for t, tower in (("q", model.query_tower), ("d", model.document_tower)):
        tower.load_state_dict({"embedding.embedding.weight": torch.tensor(g[f"init_{t}_embedding"]),
                               "feed_forward.0.weight": torch.tensor(g[f"init_{t}_w1"]), "feed_forward.0.bias": torch.tensor(g[f"init_{t}_b1"]),
                               "feed_forward.2.weight": torch.tensor(g[f"init_{t}_w2"]), "feed_forward.2.bias": torch.tensor(g[f"init_{t}_b2"])})
    model = model.to(DEV)
    B, L = g["step0_q_ids"].shape
    tr = tt.FusedTrainer(model, loss="in_batch", temperature=0.1, lr=1e-3, batch_size=B, max_len=L,
                         precision="fp32", use_cuda_graph=graph)
    for step in range(3):
        loss = tr.step(torch.tensor(g[f"step{step}_q_ids"]), torch.tensor(g[f"step{step}_d_ids"]))
        close(loss, g[f"step{step}_loss"], rtol=2e-5)
        for t, tower in (("q", model.query_tower), ("d", model.document_tower)):
            sd = tower.state_dict()
            for k, key in (("embedding", "embedding.embedding.weight"), ("w1", "feed_forward.0.weight"),
                           ("b1", "feed_forward.0.bias"), ("w2", "feed_forward.2.weight"), ("b2", "feed_forward.2.bias")):
                close(sd[key], g[f"step{step}_{t}_{k}"], rtol=2e-5)


def test_reference_style_loop_with_torch_optimizer(golden_dir):
    """The reference's own loop shape: model(q,p,n); loss_fn; zero_grad; backward; AdamW.step (train.py:120-139)."""
    import two_towers_b200 as tt
    g = np.load(os.path.join(golden_dir, "train_triplet_3steps.npz"))
    V, E = g["init_embedding"].shape; H = g["init_w1"].shape[0]
    emb = tt.embeddings.build("lookup", V, embedding_dim=E)
    model = tt.build_two_tower("mean", emb, hidden_dim=H, tied_weights=True)
    model.query_tower.load_state_dict({"embedding.embedding.weight": torch.tensor(g["init_embedding"]),
                                       "feed_forward.0.weight": torch.tensor(g["init_w1"]), "feed_forward.0.bias": torch.tensor(g["init_b1"]),
                                       "feed_forward.2.weight": torch.tensor(g["init_w2"]), "feed_forward.2.bias": torch.tensor(g["init_b2"])})
    model = model.to(DEV)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    loss_fn = tt.losses.build("triplet", margin=0.2)
    for step in range(3):
        qv, dv, nv = model(cu(g[f"step{step}_q_ids"]), cu(g[f"step{step}_d_ids"]), cu(g[f"step{step}_n_ids"]))
        loss = loss_fn(qv, dv, nv)
        opt.zero_grad(); loss.backward(); opt.step()
        close(loss, g[f"step{step}_loss"], rtol=2e-5)
        close(model.query_tower.state_dict()["feed_forward.2.weight"], g[f"step{step}_w2"], rtol=2e-5)
        close(model.query_tower.state_dict()["embedding.embedding.weight"], g[f"step{step}_embedding"], rtol=2e-5)


def test_fused_trainer_full_shape_runs_and_learns():
    """B=4096, L=64, E=64, H=256 (BASELINE config 1): the fused step lowers the loss on a fixed batch."""
    import two_towers_b200 as tt
    torch.manual_seed(0)
    emb = tt.embeddings.build("lookup", 128, embedding_dim=64)
    model = tt.build_two_tower("mean", emb, hidden_dim=256, tied_weights=True).to(DEV)
    rng = np.random.default_rng(0)
    q = torch.tensor(make_ids(rng, 4096, 64, 128)); d = torch.tensor(make_ids(rng, 4096, 64, 128))
    tr = tt.FusedTrainer(model, loss="in_batch", temperature=0.1, lr=1e-3, batch_size=4096, max_len=64, precision="fp32")
    first = tr.step(q, d).item()
    for _ in range(20):
        last = tr.step(q, d).item()
    assert np.isfinite(first) and np.isfinite(last) and last < first
    assert abs(first - np.log(4096)) < 1.5
