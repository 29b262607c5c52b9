"""GPU tests of the round-2 additions (run with ``-m gpu`` on a B200):

* word-tower shapes (BASELINE configs[2], configs/word2vec_skipgram.yml:13,23,30: V = 400 000, E = 300, L = 32) on the
  bf16 tensor-core path -- E is not a multiple of 8, operands travel as pad8-pitch bf16 rows;
* avg_pool projection on the tcgen05 tensor cores (encoders.py:100-104,144-150) and Dropout in train mode, checked with
  the kernel's own counter-based keep-mask applied to the oracle;
* token ids outside the table raise IndexError (nn.Embedding semantics, embeddings.py:33-40);
* FusedTrainer checkpoint / resume (optimizer state in torch.optim.AdamW layout) and external weight loads;
* msmarco shape (configs/msmarco_gpu.yml:24-31: untied, H = 128) with 4x repeated positives, multiple_negatives N = 4.
"""
import copy

import numpy as np
import pytest
import torch

from oracle import two_tower_oracle as O
from _parity import BF16_RTOL, check, oracle_step, tower_grads, tower_params, trainer_gates

pytestmark = pytest.mark.gpu
DEV = "cuda"


def zipf_ids(rng, B, L, V, min_len=4):
    """lengths ~U{4..L}, ids Zipf(1.07) clipped to [2, V), 1 = UNK with p = 0.02, zero padded."""
    ids = np.zeros((B, L), np.int64)
    lens = rng.integers(min_len, L + 1, B)
    x = np.minimum(rng.zipf(1.07, (B, L)) + 1, V - 1)
    x[rng.random((B, L)) < 0.02] = 1
    mask = np.arange(L)[None, :] < lens[:, None]
    ids[mask] = x[mask]
    return ids


# ------------------------------------------------------------------------------------------
# word tower: E = 300
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("R,E,H", [(1000, 300, 256), (8192, 300, 256), (257, 50, 64), (512, 100, 128)])
def test_mlp_bf16_any_embedding_dim(R, E, H):
    """TT_PREC_BF16 tower MLP with E % 8 != 0 (GloVe 50 / 100 / word2vec 300): fp64 oracle, kernel's own gates."""
    import two_towers_b200 as tt
    rng = np.random.default_rng(R + E)
    x = rng.standard_normal((R, E)).astype(np.float32) * 0.3
    w1 = (rng.standard_normal((H, E)) / np.sqrt(E)).astype(np.float32)
    b1 = (rng.standard_normal(H) * 0.1).astype(np.float32)
    w2 = (rng.standard_normal((H, H)) / np.sqrt(H)).astype(np.float32)
    b2 = (rng.standard_normal(H) * 0.1).astype(np.float32)
    dy = rng.standard_normal((R, H)).astype(np.float32)
    t = lambda a: torch.tensor(a, device=DEV)
    y, h1, z, yb = tt.ops.mlp_fwd(t(x), t(w1), t(b1), t(w2), t(b2), precision="bf16", want_bf16=True)
    f = np.float64
    a1 = x.astype(f) @ w1.astype(f).T + b1
    rz = np.maximum(a1, 0) @ w2.astype(f).T + b2
    print(f"  bf16 MLP R={R} E={E} H={H}")
    check(z, rz, BF16_RTOL, "z"); check(y, O.normalize(rz), BF16_RTOL, "y")
    dx, dw1, db1, dw2, db2 = tt.ops.mlp_bwd(t(dy), t(x), t(w1), t(w2), h1, z, True, precision="bf16")
    h1v = h1.view(torch.bfloat16).reshape(-1)[:R * H].reshape(R, H)
    mask = h1v.float().cpu().numpy() > 0
    assert (mask != (a1 > 0)).mean() < 0.02
    dz = O.normalize_bwd(dy.astype(f), rz)
    da1 = (dz @ w2.astype(f)) * mask
    check(dw2, dz.T @ np.maximum(a1, 0), BF16_RTOL, "dw2"); check(db2, dz.sum(0), BF16_RTOL, "db2")
    check(dw1, da1.T @ x.astype(f), BF16_RTOL, "dw1"); check(db1, da1.sum(0), BF16_RTOL, "db1")
    check(dx, da1 @ w1.astype(f), BF16_RTOL, "dx")
    dx2, dw1b, _, _, _ = tt.ops.mlp_bwd(t(dy), t(x), t(w1), t(w2), h1, z, True, precision="bf16")
    assert torch.equal(dx, dx2) and torch.equal(dw1, dw1b)


@pytest.mark.parametrize("loss,trainable,V", [("in_batch", True, 400_000), ("triplet", True, 50_000), ("in_batch", False, 400_000)])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_word_tower_trainer_step(loss, trainable, V, precision):
    """C3a (frozen table, configs/word2vec_skipgram.yml:25 trainable: false) and C3b (trainable: sorted-segment
    scatter-add + dense AdamW over the table) at V = 400 000, E = 300, L = 32 with Zipf ids: loss and every gradient
    against the fp64 oracle of the whole step; the large-V embedding backward is bitwise repeatable."""
    import two_towers_b200 as tt
    E, H, L, B = 300, 256, 32, 512
    rng = np.random.default_rng(7)
    torch.manual_seed(0)
    emb = tt.embeddings.build("lookup", V, embedding_dim=E)
    emb.embedding.weight.requires_grad_(trainable)
    model = tt.build_two_tower("mean", emb, hidden_dim=H, tied_weights=True).to(DEV)
    ids = [torch.tensor(zipf_ids(rng, B, L, V)) for _ in range(3)]
    pq0 = tower_params(model.query_tower)
    tr = tt.FusedTrainer(model, loss=loss, temperature=0.1, margin=2.5, batch_size=B, max_len=L, precision=precision,
                         use_cuda_graph=False)
    assert tr.train_table == trainable and not tr.embed_fused
    got_loss = tr.step(*ids[:tr.passes]).item()
    gates = trainer_gates(tr) if precision == "bf16" else None
    rl, gq, _, flips = oracle_step(pq0, pq0, ids[0].numpy(), ids[1].numpy(), ids[2].numpy() if loss == "triplet" else None,
                                   loss=loss, temperature=0.1, margin=2.5, gates=gates)
    tol = BF16_RTOL if precision == "bf16" else 1e-4
    print(f"  word tower {loss} trainable={trainable} {precision}: loss {got_loss:.6f} oracle {rl:.6f} gates flipped {flips:.2%}")
    assert abs(got_loss - rl) <= tol * abs(rl) and flips < 0.01
    named = {"w1": model.query_tower.feed_forward[0].weight, "b1": model.query_tower.feed_forward[0].bias,
             "w2": model.query_tower.feed_forward[2].weight, "b2": model.query_tower.feed_forward[2].bias}
    if trainable:
        named["embedding"] = emb.embedding.weight
    for k, p in named.items():
        check(p.grad, gq[k], tol, f"grad {k}")
    if trainable:
        assert float(emb.embedding.weight.grad[0].abs().max()) == 0.0     # padding row
    tr.check()


def test_bad_token_id_raises_index_error():
    """nn.Embedding raises IndexError for ids outside [0, V) (embeddings.py:33-40): here the kernel records the id and
    the host raises at its next synchronisation point instead of silently masking it."""
    import two_towers_b200 as tt
    table = torch.randn(50, 16, device=DEV)
    ids = torch.randint(1, 50, (8, 12), device=DEV)
    tt.ops.embed_pool_fwd(ids, table)
    torch.cuda.synchronize()
    tt._lib.raise_on_bad_ids()                                 # clean batch: nothing recorded
    ids[3, 4] = 50
    pooled, _, _ = tt.ops.embed_pool_fwd(ids, table)
    torch.cuda.synchronize()
    with pytest.raises(IndexError, match="token id 50"):
        tt._lib.raise_on_bad_ids()
    tt._lib.raise_on_bad_ids()                                 # the record was cleared
    assert torch.isfinite(pooled).all()                        # no out-of-bounds read happened
    ids[3, 4] = -7
    tt.ops.embed_gather(ids, table)
    torch.cuda.synchronize()
    with pytest.raises(IndexError, match="-7"):
        tt._lib.raise_on_bad_ids()
    # through the trainer: the asynchronous loss read surfaces it
    emb = tt.embeddings.build("lookup", 64, embedding_dim=64)
    model = tt.build_two_tower("mean", emb, hidden_dim=64, tied_weights=True).to(DEV)
    tr = tt.FusedTrainer(model, loss="in_batch", batch_size=128, max_len=16, precision="bf16")
    q = torch.randint(1, 64, (128, 16)); d = torch.randint(1, 64, (128, 16))
    tr.step(q, d); tr.read_loss_async()()
    q[5, 5] = 64
    tr.step(q, d)
    with pytest.raises(IndexError):
        tr.read_loss_async()()


# ------------------------------------------------------------------------------------------
# avg_pool tower: tensor-core projection, dropout
# ------------------------------------------------------------------------------------------
def _avg_case(R, E, H, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((R, E)) * 0.5
    w = rng.standard_normal((H, E)) / np.sqrt(E)
    b = rng.standard_normal(H) * 0.1
    gamma = 1.0 + 0.1 * rng.standard_normal(H)
    beta = 0.1 * rng.standard_normal(H)
    dy = rng.standard_normal((R, H))
    return x, w, b, gamma, beta, dy


def _avg_oracle(x, w, b, gamma, beta, dy, keep=None, p=0.0):
    """encoders.py:144-150 in fp64: Linear -> Dropout (given keep-mask, kept values / (1 - p)) -> LayerNorm -> normalise."""
    a = x @ w.T + b
    if keep is not None:
        a = np.where(keep, a / (1.0 - p), 0.0)
    ln, xhat, rstd = O.layer_norm(a, gamma, beta)
    y = O.normalize(ln)
    dz = O.normalize_bwd(dy, ln)
    da, dgamma, dbeta = O.layer_norm_bwd(dz, xhat, rstd, gamma)
    if keep is not None:
        da = np.where(keep, da / (1.0 - p), 0.0)
    return y, dict(dx=da @ w, dw=da.T @ x, db=da.sum(0), dgamma=dgamma, dbeta=dbeta)


@pytest.mark.parametrize("R,E,H", [(4096, 64, 256), (1000, 300, 128), (130, 48, 96)])
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("bf16", BF16_RTOL)])
@pytest.mark.parametrize("drop", [0.0, 0.1])
def test_proj_ln_precision_and_dropout(R, E, H, precision, tol, drop):
    import two_towers_b200 as tt
    x, w, b, gamma, beta, dy = _avg_case(R, E, H, R + E + H)
    t = lambda a: torch.tensor(a, dtype=torch.float32, device=DEV)
    seed = 0xC0FFEE + R
    step = torch.tensor([5, 0], dtype=torch.int64, device=DEV)
    keep = tt.ops.dropout_keep_mask(seed, R, H, drop, seed_step=5) if drop > 0 else None
    if keep is not None:
        assert abs(keep.mean() - (1 - drop)) < 0.01
    ry, rg = _avg_oracle(x, w, b, gamma, beta, dy, keep, drop)
    tx, tw, tb, tg, tbt, tdy = map(t, (x, w, b, gamma, beta, dy))
    y, a, stats, z = tt.ops.proj_ln_fwd(tx, tw, tb, tg, tbt, True, drop, drop > 0, seed, precision=precision, seed_step=step)
    print(f"  proj+LN R={R} E={E} H={H} {precision} dropout={drop}")
    check(y, ry, tol, "y")
    dx, dw, db, dg, dbt = tt.ops.proj_ln_bwd(tdy, tx, tw, tg, a, stats, z, True, drop, drop > 0, seed, precision=precision,
                                             seed_step=step)
    for k, v in (("dx", dx), ("dw", dw), ("db", db), ("dgamma", dg), ("dbeta", dbt)):
        check(v, rg[k], tol, k)
    if drop > 0:                                               # a different step counter draws a different mask
        step2 = torch.tensor([6, 0], dtype=torch.int64, device=DEV)
        y2, *_ = tt.ops.proj_ln_fwd(tx, tw, tb, tg, tbt, True, drop, True, seed, precision=precision, seed_step=step2)
        assert not torch.equal(y, y2)
        y3, *_ = tt.ops.proj_ln_fwd(tx, tw, tb, tg, tbt, True, drop, False, seed, precision=precision, seed_step=step2)
        check(y3, _avg_oracle(x, w, b, gamma, beta, dy)[0], tol, "y (eval mode: dropout off)")


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", BF16_RTOL)])
def test_fused_trainer_avg_pool_applies_dropout(precision, tol):
    """FusedTrainer on AveragePoolingTower(dropout=0.1) in train mode (the reference loop calls model.train(),
    train.py:86): the fused step must apply Dropout, with a fresh mask per optimizer step under CUDA-graph replay."""
    import two_towers_b200 as tt
    torch.manual_seed(3)
    V, E, H, B, L = 128, 64, 128, 256, 32
    emb = tt.embeddings.build("lookup", V, embedding_dim=E)
    model = tt.build_two_tower("avg_pool", emb, hidden_dim=H, tied_weights=True, dropout=0.1).to(DEV)
    model.train()
    g = torch.Generator().manual_seed(5)
    q, d, n = (torch.randint(0, V, (B, L), generator=g) for _ in range(3))
    tw = model.query_tower
    p0 = {k: v.detach().double().cpu().numpy() for k, v in
          dict(embedding=tw.embedding.embedding.weight, w=tw.projection[0].weight, b=tw.projection[0].bias,
               gamma=tw.projection[2].weight, beta=tw.projection[2].bias).items()}
    tr = tt.FusedTrainer(model, loss="triplet", margin=2.5, batch_size=B, max_len=L, precision=precision, use_cuda_graph=True)
    l0 = tr.step(q, d, n).item()                               # optimizer step 0 (seed_step = 0)
    a_after_0 = tr.saved[0]["a"].clone()
    # oracle with the kernel's mask: one tied tower over the stacked [q|d|n] rows (group 0, rank 0)
    _, train, seed = tr._dropout_cfg(tw, 0)
    assert train == 1
    keep = tt.ops.dropout_keep_mask(seed, 3 * B, H, 0.1, seed_step=0)
    ids = np.concatenate([q.numpy(), d.numpy(), n.numpy()])
    pooled, _ = O.masked_mean_pool(ids, p0["embedding"], np.float64)
    y, _ = _avg_oracle(pooled, p0["w"], p0["b"], p0["gamma"], p0["beta"], np.zeros((3 * B, H)), keep, 0.1)
    rl = O.triplet_loss(y[:B], y[B:2 * B], y[2 * B:], 2.5)
    print(f"  avg_pool trainer {precision}: loss {l0:.6f} oracle (same dropout mask) {rl:.6f}")
    assert abs(l0 - rl) <= tol * abs(rl)
    dropped = (a_after_0 == 0).float().mean().item()
    assert 0.07 < dropped < 0.13                               # ~10 % of the projection outputs were zeroed
    tr.step(q, d, n)                                           # graph replay, optimizer step 1: a new mask
    z1 = (tr.saved[0]["a"] == 0)
    assert (z1 != (a_after_0 == 0)).float().mean().item() > 0.1
    model.eval()
    tr_eval = tt.FusedTrainer(copy.deepcopy(model), loss="triplet", margin=2.5, batch_size=B, max_len=L, precision=precision,
                              use_cuda_graph=False)
    tr_eval.step(q, d, n)
    assert (tr_eval.saved[0]["a"] == 0).float().mean().item() < 0.01


# ------------------------------------------------------------------------------------------
# trainer state
# ------------------------------------------------------------------------------------------
def test_fused_trainer_checkpoint_resume_and_external_loads():
    """save_checkpoint (twotower/utils.py) stores model.state_dict() + optimizer.state_dict(): a run resumed from them
    continues bit-identically; the optimizer state loads into torch.optim.AdamW; external weight loads reach the bf16
    shadow after sync_from_model()."""
    import two_towers_b200 as tt
    torch.manual_seed(0)
    V, E, H, B, L = 128, 64, 256, 512, 32
    def fresh():
        torch.manual_seed(0)
        emb = tt.embeddings.build("lookup", V, embedding_dim=E)
        return tt.build_two_tower("mean", emb, hidden_dim=H, tied_weights=True).to(DEV)
    g = torch.Generator().manual_seed(9)
    batches = [(torch.randint(0, V, (B, L), generator=g), torch.randint(0, V, (B, L), generator=g)) for _ in range(6)]
    m1 = fresh()
    t1 = tt.FusedTrainer(m1, loss="in_batch", batch_size=B, max_len=L, precision="bf16")
    for q, d in batches[:3]:
        t1.step(q, d)
    ck_model = {k: v.clone() for k, v in m1.state_dict().items()}
    ck_opt = t1.state_dict()
    ref_losses = [t1.step(q, d).item() for q, d in batches[3:]]
    # resume in a fresh trainer
    m2 = fresh()
    t2 = tt.FusedTrainer(m2, loss="in_batch", batch_size=B, max_len=L, precision="bf16")
    m2.load_state_dict(ck_model)
    t2.load_state_dict(ck_opt)
    assert int(t2.step_count[0].item()) == 3
    got = [t2.step(q, d).item() for q, d in batches[3:]]
    assert got == ref_losses, (got, ref_losses)
    assert torch.equal(t1.flat, t2.flat)
    # torch.optim.AdamW accepts the state and takes the same next step as the fused kernel (fp32 parameters)
    m3 = fresh(); m3.load_state_dict(ck_model)
    opt = torch.optim.AdamW([p for p in m3.parameters() if p.requires_grad], lr=1e-3)
    opt.load_state_dict(ck_opt)
    assert opt.state_dict()["state"][0]["exp_avg"].shape == ck_opt["state"][0]["exp_avg"].shape
    # without sync_from_model the bf16 shadow is stale: kernels_per_step leaves the state untouched
    before = t2.flat.clone(); cnt = int(t2.step_count[0].item())
    t2.load_batch(*batches[0])
    assert t2.kernels_per_step() > 0
    assert torch.equal(before, t2.flat) and int(t2.step_count[0].item()) == cnt
    with torch.no_grad():
        m2.query_tower.feed_forward[2].bias.add_(1.0)
    t2.sync_from_model()
    off = t2.offsets[id(m2.query_tower.feed_forward[2].bias)]
    assert torch.equal(t2.flat_bf16[off:off + H].float(), t2.flat[off:off + H].bfloat16().float())


# ------------------------------------------------------------------------------------------
# msmarco shape (C4)
# ------------------------------------------------------------------------------------------
def test_msmarco_shape_repeated_positives_and_multiple_negatives():
    """configs/msmarco_gpu.yml:24-31 (untied mean towers, E = 64, H = 128) on a batch where every positive appears 4 times
    (presets/multi_pos_multi_neg.yml:12 negatives_per_pos: 4 -> duplicated (q, d+) rows; the reference applies no
    false-negative masking): in-batch step vs the fp64 oracle, and multiple_negatives_loss with N = 4 (losses.py:47-85).

    Conditioning.  With 64 random characters per row every mean-pooled input -- hence every hidden activation h1_r -- is
    nearly the same vector at initialisation, and sum_r dz_r = 0 for a uniform softmax, so the weight gradients
    dW = sum_r dz_r (x) h1_r are small residues of cancelling terms: the bf16 rounding of the GEMM operands (2^-9 per
    element, the definition of TT_PREC_BF16) is amplified by |h_mean| / |h_r - h_mean| and reaches 2-6 % there (measured,
    gpurun_out r02c; the oracle on bf16-rounded tower outputs alone moves the gradients by only 0.4 %).  Rows of 1-6
    tokens keep the activations apart; on them every gradient holds the stated 2e-2 (document-tower db2: see
    test_fused_trainer_bf16_tracks_fp32)."""
    import two_towers_b200 as tt
    torch.manual_seed(0)
    V, E, H, B, L = 128, 64, 128, 1024, 64
    emb = tt.embeddings.build("lookup", V, embedding_dim=E)
    model = tt.build_two_tower("mean", emb, hidden_dim=H, tied_weights=False).to(DEV)
    g = torch.Generator().manual_seed(2)

    def ids():
        x = torch.randint(1, V, (B // 4, L), generator=g)
        lens = torch.randint(1, 7, (B // 4, 1), generator=g)
        return torch.where(torch.arange(L)[None, :] < lens, x, torch.zeros_like(x)).repeat_interleave(4, 0)
    q, d = ids(), ids()
    pq0, pd0 = tower_params(model.query_tower), tower_params(model.document_tower)
    tr = tt.FusedTrainer(model, loss="in_batch", temperature=0.1, batch_size=B, max_len=L, precision="bf16")
    assert tr.local_fast and not tr.tied
    got = tr.step(q, d).item()
    rl, gq, gd, flips = oracle_step(pq0, pd0, q.numpy(), d.numpy(), loss="in_batch", temperature=0.1, gates=trainer_gates(tr))
    print(f"  msmarco shape, 4x repeated positives, bf16: loss {got:.6f} oracle {rl:.6f} (>= log 4 = {np.log(4):.4f}), gates flipped {flips:.2%}")
    assert abs(got - rl) <= BF16_RTOL * abs(rl) and rl > np.log(4) - 1e-6 and flips < 0.01
    got_q, got_d = tower_grads(model.query_tower), tower_grads(model.document_tower)
    for k in gq:
        check(got_q[k], gq[k], BF16_RTOL, f"grad query/{k}")
    for k in ("w1", "b1", "w2", "b2"):
        check(got_d[k], gd[k], 5e-2 if k == "b2" else BF16_RTOL, f"grad document/{k}")
    # multiple negatives, N = 4, on tower outputs of that shape (fp32 row kernels), B = 4096
    rng = np.random.default_rng(1)
    Bm, N = 4096, 4
    qv = O.normalize(rng.standard_normal((Bm, H))).astype(np.float32)
    pv = O.normalize(rng.standard_normal((Bm, H)) + qv).astype(np.float32)
    nv = O.normalize(rng.standard_normal((Bm, N, H))).astype(np.float32)
    t = lambda a: torch.tensor(a, device=DEV)
    loss, probs = tt.ops.multineg_fwd(t(qv), t(pv), t(nv), 0.1)
    f = np.float64
    rl = O.multiple_negatives_loss(qv.astype(f), pv.astype(f), nv.astype(f), 0.1)
    assert abs(loss.item() - rl) <= 1e-5 * abs(rl)
    dq, dp, dn = tt.ops.multineg_bwd(t(qv), t(pv), t(nv), probs, 0.1)
    rdq, rdp, rdn = O.multiple_negatives_loss_bwd(qv.astype(f), pv.astype(f), nv.astype(f), 0.1)
    check(dq, rdq, 2e-5, "multineg dq"); check(dp, rdp, 2e-5, "multineg dp"); check(dn, rdn, 2e-5, "multineg dnegs")


# ------------------------------------------------------------------------------------------
# batched search on the tensor cores
# ------------------------------------------------------------------------------------------
def _ref_topk(index_bf16, queries, k, inv=None):
    D = index_bf16.float().double().cpu().numpy()
    q = queries.double().cpu().numpy()
    sc = q @ D.T
    if inv is not None:
        sc = sc * inv.double().cpu().numpy()[None, :] / np.maximum(np.linalg.norm(q, axis=1, keepdims=True), 1e-8)
    rs, ri = O.topk_lower_index(sc, k)
    return sc, rs, ri


@pytest.mark.parametrize("N,H,k,nq", [(200_003, 256, 100, 16), (200_003, 256, 100, 130), (50_000, 128, 10, 5), (300, 64, 100, 7),
                                      (100, 256, 100, 4), (500_000, 256, 100, 64), (4099, 192, 128, 33)])
def test_topk_scan_batched_matches_exact_topk(N, H, k, nq):
    """tt_topk_scan_batched (tcgen05 Q D^T + fused top-k) vs torch.topk semantics on fp64 scores of the same bf16 index
    (ids identical except at ties within tolerance, ties -> lower index) and vs the single-query scan kernel."""
    import two_towers_b200 as tt
    g = torch.Generator(device=DEV).manual_seed(N + nq)
    D = tt.ops.cast_bf16(torch.nn.functional.normalize(torch.randn(N, H, device=DEV, generator=g), dim=-1))
    q = torch.nn.functional.normalize(torch.randn(nq, H, device=DEV, generator=g), dim=-1)
    if N > 70_000:
        D[70_000] = D[3]                                       # exact duplicate far away: tie -> lower id first
        q[0] = D[3].float()
    assert tt.ops.topk_scan_batched_ok(D, k)
    s, i = tt.ops.topk_scan_batched(D, q, k, id_offset=1000)
    sc, rs, ri = _ref_topk(D, q, k)
    assert O.topk_ids_match(i.cpu().numpy() - 1000, s.cpu().numpy(), ri, rs, sc, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(s.cpu().numpy(), rs, rtol=2e-5, atol=2e-6)
    if N > 70_000:
        assert i[0, :2].tolist() == [1003, 71_000]
    s1, i1 = tt.ops.topk_scan(D, q, k, cosine=False, id_offset=1000)      # the CUDA-core kernel on the same index
    assert int((i != i1).sum()) <= max(2, int(0.002 * i.numel()))   # identical up to last-ulp score differences at the boundary
    s2, i2 = tt.ops.topk_scan_batched(D, q, k, id_offset=1000)
    assert torch.equal(i, i2) and torch.equal(s, s2)           # deterministic


def test_topk_scan_batched_cosine_ties_and_api():
    import two_towers_b200 as tt
    g = torch.Generator(device=DEV).manual_seed(5)
    N, H, k = 30_000, 256, 50
    D = tt.ops.cast_bf16(torch.randn(N, H, device=DEV, generator=g) * (0.5 + torch.rand(N, 1, device=DEV, generator=g)))   # not unit rows
    q = torch.randn(9, H, device=DEV, generator=g) * 3.0
    inv = tt.ops.index_row_inv_norms(D)
    s, i = tt.ops.topk_scan_batched(D, q, k, row_inv_norms=inv)
    sc, rs, ri = _ref_topk(D, q, k, inv)
    assert O.topk_ids_match(i.cpu().numpy(), s.cpu().numpy(), ri, rs, sc, rtol=1e-5, atol=1e-6)
    s1, i1 = tt.ops.topk_scan(D, q, k, cosine=True)
    assert int((i != i1).sum()) <= 2
    # every document identical: all scores tie, the answer is rows 0 .. k-1 in order
    Dsame = D[:1].expand(5000, H).contiguous()
    s, i = tt.ops.topk_scan_batched(Dsame, q, k)
    assert torch.equal(i, torch.arange(k, device=DEV).expand(9, k))
    # through TwoTowerSearch.search_batch: same documents as per-query search on the bf16 index
    torch.manual_seed(0)
    tok = tt.CharTokeniser().fit(["abcdefghijklmnopqrstuvwxyz 0123456789"])
    emb = tt.embeddings.build("lookup", tok.vocab_size, embedding_dim=64)
    model = tt.build_two_tower("mean", emb, hidden_dim=128, tied_weights=True).to(DEV)
    docs = [f"document number {i} about topic {i % 17} and {(i * 7) % 13}" for i in range(3000)]
    srch = tt.TwoTowerSearch(model, tok, device=DEV, index_dtype="bf16")
    srch.index_documents(docs)
    queries = [f"topic {j} and {j % 13}" for j in range(12)]
    batched = srch.search_batch(queries, top_k=20)
    srch.batched_min_queries = 10 ** 9                          # force the per-query kernel
    single = srch.search_batch(queries, top_k=20)
    for a, b in zip(batched, single):
        assert [r["document"] for r in a] == [r["document"] for r in b]
        np.testing.assert_allclose([r["score"] for r in a], [r["score"] for r in b], rtol=1e-4)
