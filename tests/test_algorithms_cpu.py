"""CPU restatements (numpy, fp64) of the two algorithmic restructurings the CUDA kernels rely on, checked against the oracle /
a plain sort.  They document WHY the kernels' results equal the reference's; the kernels themselves are checked on the GPU
(tests/test_gpu_onepass.py, tests/test_gpu_parity.py).

1. Stored-E form of the in-batch softmax gradients (tc_inbatch.cu, MODE 2 + MODE 3; reference twotower/losses.py:107-116):
   with a fixed shift m >= every logit, E = exp(logit - m) with the positives left out, L_i = sum_j E_ij + E_pos(i):
       dq_i = c (sum_j E_ij d_j / L_i - (1 - P_pos(i)) d_pos(i)),    dd_j = c (sum_i E_ij q_i / L_i - (1 - P_pos(i(j))) q_i(j)),
   c = grad / (B t), 1 - P_pos = (L - E_pos) / L -- S is formed once and no exponential is taken twice.
2. Top-k of G descending lists from their heads (topk_scan.cu, select_from_lists / cta_select_runs): the k-th largest of the
   first ceil(k / G) keys of every list is a lower bound of the k-th largest key overall, so selecting among the keys that
   reach it is exact.
"""
import numpy as np
import pytest

from oracle import two_tower_oracle as O


@pytest.mark.parametrize("Bq,Bd,off,temp", [(64, 64, 0, 0.1), (37, 101, 13, 0.05), (128, 300, 172, 1.0)])
def test_stored_e_decomposition_equals_oracle_gradients(Bq, Bd, off, temp):
    rng = np.random.default_rng(Bq + Bd)
    q = O.normalize(rng.standard_normal((Bq, 48)))
    d = O.normalize(rng.standard_normal((Bd, 48)))
    idx = np.arange(Bq)
    d[idx + off] = O.normalize(d[idx + off] + 2.0 * q)           # peaked softmax rows: the cancellation-free form matters
    m = 1.0 / temp                                               # unit rows: |logit| <= 1 / temperature
    logits = (q @ d.T) / temp
    assert logits.max() <= m + 1e-12
    E = np.exp(logits - m)
    e_pos = E[idx, idx + off].copy()
    E[idx, idx + off] = 0.0                                      # the positives stay out of the stored tiles and of the sums
    L_off = E.sum(1)
    L = L_off + e_pos
    w_pos = L_off / L                                            # 1 - P_pos, free of cancellation
    c = 0.5 / (Bq * temp)                                        # grad = 0.5
    dq = c * ((E @ d) / L[:, None] - w_pos[:, None] * d[idx + off])
    dd = c * (E.T @ (q / L[:, None]))
    dd[idx + off] -= c * w_pos[:, None] * q                      # rank-one terms of the positives
    lse = m + np.log(L)
    loss = (lse - logits[idx, idx + off]).mean()
    rl, rlse = O.in_batch_loss(q, d, temp, off)
    rdq, rdd = O.in_batch_loss_bwd(q, d, temp, off, grad=0.5)
    assert abs(loss - rl) <= 1e-12 * max(1.0, abs(rl))
    np.testing.assert_allclose(lse, rlse, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(dq, rdq, rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(dd, rdd, rtol=1e-10, atol=1e-14)


def _select_from_lists(lists, k):
    """numpy restatement of select_from_lists: `lists` [G, n] descending, unique keys (0 = padding)."""
    G = lists.shape[0]
    r = -(-k // G)
    heads = lists[:, :r].ravel()
    live = np.sort(heads[heads != 0])[::-1]
    if live.size < k:
        return None                                             # fewer than k live heads: the kernel falls back to the radix select
    thr = live[k - 1]
    surv = lists[(lists >= thr) & (lists != 0)]
    return np.sort(surv)[::-1][:k]


@pytest.mark.parametrize("G,n,k", [(148, 100, 100), (8, 100, 100), (16, 64, 100), (64, 64, 128), (3, 10, 7), (5, 4, 1)])
def test_topk_from_list_heads_is_exact(G, n, k):
    rng = np.random.default_rng(G * 1000 + k)
    for trial in range(20):
        keys = rng.permutation(np.arange(1, 10 * G * n + 1, dtype=np.int64))[:G * n].reshape(G, n)
        if trial % 4 == 1:                                       # one list holds all the large keys
            keys = np.sort(keys.ravel())[::-1].reshape(G, n)
        if trial % 4 == 2:                                       # short lists: zero padding at the end
            for g in range(G):
                keys[g, rng.integers(0, n + 1):] = 0
        lists = -np.sort(-keys, axis=1)
        got = _select_from_lists(lists, k)
        want = np.sort(lists[lists != 0])[::-1][:k]
        if got is None:
            heads = lists[:, :-(-k // G)]
            assert (heads != 0).sum() < k
            continue
        np.testing.assert_array_equal(got, want)


def test_bitonic_register_layout_sorts():
    """warp_sort_desc: 32 * KPL keys, element e = j * 32 + lane; distances below 32 exchange between lanes (shuffles), the
    others between registers of one lane -- the compare directions below are the kernel's."""
    rng = np.random.default_rng(3)
    for kpl in (2, 4, 8):
        M = 32 * kpl
        for _ in range(10):
            v = rng.permutation(M).astype(np.int64).reshape(kpl, 32)          # v[j][lane]
            size = 2
            while size <= M:
                stride = size >> 1
                while stride > 0:
                    nv = v.copy()
                    for j in range(kpl):
                        for lane in range(32):
                            e = (j << 5) | lane
                            desc = (e & size) == 0
                            if stride >= 32:
                                sj = stride >> 5
                                if (j & sj) == 0:
                                    a, b = v[j][lane], v[j | sj][lane]
                                    if (a < b) == desc:
                                        nv[j][lane], nv[j | sj][lane] = b, a
                            else:
                                other = v[j][lane ^ stride]
                                keep_max = ((lane & stride) == 0) == desc
                                nv[j][lane] = max(v[j][lane], other) if keep_max else min(v[j][lane], other)
                    v = nv
                    stride >>= 1
                size <<= 1
            flat = v.reshape(-1)                                               # e = j * 32 + lane
            assert np.array_equal(flat, np.sort(flat)[::-1])
