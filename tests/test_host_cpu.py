"""CPU-side tests: host logic, tokenisers vs golden, registries, C-ABI export surface, and the
loud failure when no B200 is present.  No GPU compute here."""
import ctypes
import json
import os
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="session", autouse=True)
def built_library():
    so = os.path.join(ROOT, "two_towers_b200", "csrc", "libtt_b200.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.dirname(so), "-j8"], stdout=subprocess.DEVNULL)
    return so


def test_cabi_exports_every_declared_symbol(built_library):
    hdr = open(os.path.join(ROOT, "include", "tt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(tt_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = ctypes.CDLL(built_library)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/tt_b200.h but not exported"
    from two_towers_b200 import _lib
    assert set(_lib.SIGNATURES) == declared       # the ctypes table mirrors the header exactly
    assert _lib.load().tt_abi_version() == _lib.TT_ABI_VERSION == 6


def test_shape_predicates_need_no_gpu(built_library):
    """The planning entry points of the stored-E loss and the fused sharded search are pure host arithmetic: which shapes take
    which path is decided (and testable) without a device."""
    import ctypes as C
    from two_towers_b200 import _lib
    lib = _lib.load()
    # stored-E loss: H >= 128 (H % 64 == 0, <= 256) and E = Bq x roundup(Bd, 64) bf16 within the L2 budget (48 MB by default)
    assert lib.tt_inbatch_ce_stash_ok(4096, 4096, 256) == 1 and lib.tt_inbatch_ce_stash_ok(4096, 4096, 128) == 1
    assert lib.tt_inbatch_ce_stash_ok(4096, 4096, 64) == 0 and lib.tt_inbatch_ce_stash_ok(4096, 4096, 200) == 0
    assert lib.tt_inbatch_ce_stash_ok(8192, 8192, 256) == 0          # 128 MB: the recomputing document pass runs
    assert lib.tt_inbatch_ce_stash_ok(4096, 32768, 256) == 0         # one rank's share of an 8-GPU global-negatives step
    assert lib.tt_inbatch_ce_stash_ok(0, 10, 256) == 0
    n = lib.tt_inbatch_ce_stash_bytes(1000, 3001, 256)               # E rows are padded to 64 columns; X / L and 1 - P_pos follow
    assert n >= 1000 * 3008 * 2 + 1000 * 256 * 2 + 1000 * 4 and n % 256 == 0
    # fused sharded search: double-buffered exchange, one CTA per query, slots that hold nq * k keys
    x = _lib.P2P()
    x.world, x.rank, x.slot_bytes, x.double_buffered, x.ctas = 8, 3, 1280, 1, 1
    assert lib.tt_topk_scan_p2p_ok(C.byref(x), 1, 100) == 1
    assert lib.tt_topk_scan_p2p_ok(C.byref(x), 2, 100) == 0          # ctas != nq
    x.ctas = 2
    assert lib.tt_topk_scan_p2p_ok(C.byref(x), 2, 100) == 0          # 2 x 100 keys do not fit a 1280-byte slot
    x.slot_bytes = 2048
    assert lib.tt_topk_scan_p2p_ok(C.byref(x), 2, 100) == 1
    x.double_buffered = 0
    assert lib.tt_topk_scan_p2p_ok(C.byref(x), 2, 100) == 0
    assert lib.tt_topk_scan_p2p_ok(None, 1, 100) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly_no_fallback():
    from two_towers_b200 import _lib, ops
    lib = _lib.load()
    buf = (ctypes.c_float * 16)()
    rc = lib.tt_embed_pool_fwd(ctypes.addressof(buf), 8, ctypes.addressof(buf), 1, 1, 2, 4,
                               ctypes.addressof(buf), ctypes.addressof(buf), None, None, None)
    assert rc == -3 and "no CPU fallback" in _lib.last_error()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.embed_pool_fwd(torch.zeros(2, 3, dtype=torch.int64), torch.zeros(4, 4))
    import two_towers_b200 as tt
    emb = tt.embeddings.build("lookup", 10, embedding_dim=4)
    model = tt.build_two_tower("mean", emb, hidden_dim=8, tied_weights=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.ones(2, 3, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tt.FusedTrainer(model, batch_size=2, max_len=3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tt.in_batch_sampled_softmax_loss(torch.randn(4, 8), torch.randn(4, 8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "two_towers_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|oracle[./]", src, flags=re.M), f"{fn} uses oracle/"


def test_tokenisers_match_reference_golden(golden_dir):
    from two_towers_b200 import tokenisers as T
    g = json.load(open(os.path.join(golden_dir, "tokenisers.json")))
    ct = T.build("char").fit(g["corpus"])
    wt = T.build("word").fit(g["corpus"])
    assert ct.string_to_index == g["char"]["vocab"] and ct.vocab_size == g["char"]["vocab_size"]
    assert wt.word_to_index == g["word"]["vocab"] and wt.vocab_size == g["word"]["vocab_size"]
    for i, p in enumerate(g["probes"]):
        assert ct.encode(p) == g["char"]["encoded"][i]
        assert ct.truncate_and_pad(ct.encode(p), 16) == g["char"]["padded16"][i]
        assert wt.encode(p) == g["word"]["encoded"][i]
        assert wt.truncate_and_pad(wt.encode(p), 8) == g["word"]["padded8"][i]
        assert wt.truncate_and_pad(wt.encode(p)) == g["word"]["padded_default"][i]
    # batch path == per-item path
    b = ct.encode_batch(g["probes"], 16).numpy()
    assert b.tolist() == g["char"]["padded16"]
    b = wt.encode_batch(g["probes"], 8).numpy()
    assert b.tolist() == g["word"]["padded8"]
    assert ct.decode(ct.encode("rockets")) == "rockets"
    with pytest.raises(ValueError):
        T.build("bpe")


def test_tokeniser_save_load_roundtrip(tmp_path):
    from two_towers_b200 import tokenisers as T
    ct = T.CharTokeniser().fit(["abc", "bcd"])
    ct.save(tmp_path / "c.pkl")
    assert T.CharTokeniser.load(tmp_path / "c.pkl").string_to_index == ct.string_to_index
    wt = T.WordTokeniser(max_len=5).fit(["a b b c", "c c c"])
    wt.save(tmp_path / "w.pkl")
    w2 = T.WordTokeniser.load(tmp_path / "w.pkl")
    assert w2.word_to_index == wt.word_to_index and w2.max_len == 5
    assert wt.word_to_index["c"] == 2 and wt.word_to_index["b"] == 3       # frequency-ranked from 2


def test_registries_and_state_dict_names():
    import two_towers_b200 as tt
    assert set(tt.embeddings.REGISTRY) == {"lookup", "word2vec", "glove"}
    assert set(tt.TOWER_REGISTRY) == {"mean", "avg_pool"}
    assert set(tt.LOSS_REGISTRY) == {"triplet", "multiple_negatives", "in_batch"}
    with pytest.raises(ValueError):
        tt.embeddings.build("nope", 10)
    with pytest.raises(ValueError):
        tt.build_tower("nope", None)
    with pytest.raises(ValueError):
        tt.losses.build("nope")
    emb = tt.embeddings.build("lookup", 11, embedding_dim=4)
    assert torch.all(emb.embedding.weight[0] == 0)                          # padding_idx row
    m = tt.build_two_tower("mean", emb, hidden_dim=8, tied_weights=False)
    keys = set(m.state_dict())
    for t in ("query_tower", "document_tower"):
        for k in ("embedding.embedding.weight", "feed_forward.0.weight", "feed_forward.0.bias",
                  "feed_forward.2.weight", "feed_forward.2.bias"):
            assert f"{t}.{k}" in keys
    assert m.query_tower.embedding is m.document_tower.embedding            # shared even when untied
    assert len(list(m.parameters())) == 9
    tied = tt.build_two_tower("mean", emb, hidden_dim=8, tied_weights=True)
    assert tied.query_tower is tied.document_tower
    a = tt.build_two_tower("avg_pool", emb, hidden_dim=8, tied_weights=True, dropout=0.0)
    assert {"query_tower.projection.0.weight", "query_tower.projection.2.weight"} <= set(a.state_dict())
    a2 = tt.build_two_tower("avg_pool", emb, hidden_dim=4, tied_weights=True)
    assert not a2.query_tower.has_projection


def test_search_api_errors_and_pickle_schema(tmp_path):
    import two_towers_b200 as tt

    class Dummy(torch.nn.Module):
        pass

    s = tt.TwoTowerSearch(Dummy(), tt.CharTokeniser(), device="cpu")
    with pytest.raises(ValueError, match="No documents indexed"):
        s.search("x")
    with pytest.raises(ValueError, match="No index to save"):
        s.save_index(tmp_path / "i.pkl")
    with pytest.raises(ValueError):
        tt.TwoTowerSearch(Dummy(), tt.CharTokeniser(), device="cpu", index_dtype="int8")
    assert issubclass(tt.TwoTowerSearch, tt.BaseSearch)
    with pytest.raises(TypeError):
        tt.BaseSearch()


def test_shard_bounds_cover_and_partition():
    from two_towers_b200.parallel import shard_bounds
    for n in (0, 1, 7, 8, 10_000_000):
        for ws in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_ranking_metrics_match_reference_golden():
    """two_towers_b200.evaluate helpers vs values produced by the reference's twotower/evaluate.py
    (tests/golden/make_golden_eval.py)."""
    import json
    from two_towers_b200 import evaluate as E
    with open(os.path.join(os.path.dirname(__file__), "golden", "eval_metrics.json")) as f:
        cases = json.load(f)
    assert len(cases) >= 100
    for c in cases:
        rel, k = c["relevance"], c["k"]
        assert E.mean_reciprocal_rank(rel) == c["mrr"]
        assert abs(E.precision_at_k(rel, k) - c["precision"]) < 1e-12
        assert abs(E.recall_at_k(rel, k, sum(rel)) - c["recall"]) < 1e-12
        assert abs(E.ndcg_at_k(rel, k) - c["ndcg"]) < 1e-12
