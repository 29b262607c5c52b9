"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE ITSELF.

Run once in the build container (the only place ``/root/reference`` exists):

    python tests/golden/make_golden.py

It imports the reference's hot-path modules by file path (SURVEY.md appendix A route (i):
skips ``twotower/__init__.py`` whose ``tools.huggingface`` import is broken on the installed
huggingface_hub), runs forward + autograd backward + AdamW on seeded inputs on CPU fp32 and
writes small ``.npz`` / ``.json`` files that are committed.  Nothing in the test-suite reads
``/root/reference`` at run time -- only these fixtures travel to the GPU box.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("TT_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def load_pkg(pkgname, root, mods):
    pkg = types.ModuleType(pkgname)
    pkg.__path__ = [root]
    sys.modules[pkgname] = pkg
    out = {}
    for m in mods:
        spec = importlib.util.spec_from_file_location(f"{pkgname}.{m}", f"{root}/{m}.py")
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"{pkgname}.{m}"] = mod
        spec.loader.exec_module(mod)
        out[m] = mod
    return out


def load_reference():
    tt = load_pkg("twotower", f"{REF}/twotower", ["tokenisers", "embeddings", "encoders", "losses"])
    inf = types.ModuleType("inference")
    inf.__path__ = [f"{REF}/inference"]
    sys.modules["inference"] = inf
    srch = load_pkg("inference.search", f"{REF}/inference/search", ["base", "two_tower"])
    return tt, srch


def npy(t):
    return t.detach().cpu().numpy().copy()


def make_ids(g, B, L, V, min_len=1, holes=True):
    ids = torch.zeros(B, L, dtype=torch.int64)
    for b in range(B):
        n = int(torch.randint(min_len, L + 1, (1,), generator=g))
        ids[b, :n] = torch.randint(1, V, (n,), generator=g)
    if holes and B >= 3:
        ids[1, 1] = 0            # mid-sequence 0 (CharTokeniser unknown char) is masked too
        ids[2, :] = 0            # all-pad row
    return ids


def tower_params(tower, kind):
    sd = {k: npy(v) for k, v in tower.state_dict().items()}
    out = {"embedding": sd["embedding.embedding.weight"]}
    if kind == "mean":
        out.update(w1=sd["feed_forward.0.weight"], b1=sd["feed_forward.0.bias"],
                   w2=sd["feed_forward.2.weight"], b2=sd["feed_forward.2.bias"])
    elif "projection.0.weight" in sd:
        out.update(w=sd["projection.0.weight"], b=sd["projection.0.bias"],
                   gamma=sd["projection.2.weight"], beta=sd["projection.2.bias"])
    return out


def grads_of(tower, kind):
    named = dict(tower.named_parameters())
    g = {"embedding": npy(named["embedding.embedding.weight"].grad)}
    if kind == "mean":
        g.update(w1=npy(named["feed_forward.0.weight"].grad), b1=npy(named["feed_forward.0.bias"].grad),
                 w2=npy(named["feed_forward.2.weight"].grad), b2=npy(named["feed_forward.2.bias"].grad))
    elif "projection.0.weight" in named:
        g.update(w=npy(named["projection.0.weight"].grad), b=npy(named["projection.0.bias"].grad),
                 gamma=npy(named["projection.2.weight"].grad), beta=npy(named["projection.2.bias"].grad))
    return g


def main():
    tt, srch = load_reference()
    tok, emb, enc, los = tt["tokenisers"], tt["embeddings"], tt["encoders"], tt["losses"]
    TwoTowerSearch = srch["two_tower"].TwoTowerSearch

    # ---------------------------------------------------------------- tokenisers
    corpus = ["how do rockets work?", "Rockets burn fuel; thrust pushes them up.",
              "what is a two tower model", "a Two-Tower model encodes queries & docs separately",
              "the the the cat sat", "", "naive cafe 123"]
    probes = ["how do towers work", "unknown ~ chars ^", "", "The cat sat on the the mat, twice!",
              "x" * 100]
    ct = tok.build("char").fit(corpus)
    wt = tok.build("word").fit(corpus)
    tok_gold = {
        "corpus": corpus, "probes": probes,
        "char": {"vocab": ct.string_to_index, "vocab_size": ct.vocab_size,
                 "encoded": [ct.encode(p) for p in probes],
                 "padded16": [ct.truncate_and_pad(ct.encode(p), 16) for p in probes]},
        "word": {"vocab": wt.word_to_index, "vocab_size": wt.vocab_size,
                 "encoded": [wt.encode(p) for p in probes],
                 "padded8": [wt.truncate_and_pad(wt.encode(p), 8) for p in probes],
                 "padded_default": [wt.truncate_and_pad(wt.encode(p)) for p in probes]},
    }
    with open(f"{OUT}/tokenisers.json", "w") as f:
        json.dump(tok_gold, f, indent=1, sort_keys=True)

    # ---------------------------------------------------------------- towers + losses
    for name, kind, V, E, H, B, L in [
        ("mean_small", "mean", 40, 16, 32, 8, 12),
        ("mean_char", "mean", 128, 64, 256, 16, 64),
        ("mean_odd", "mean", 57, 20, 24, 5, 7),
        ("avg_proj", "avg_pool", 40, 16, 32, 8, 12),
        ("avg_noproj", "avg_pool", 40, 32, 32, 8, 12),
    ]:
        torch.manual_seed(0)
        g = torch.Generator().manual_seed(1234)
        embedding = emb.build("lookup", V, embedding_dim=E)
        kw = {} if kind == "mean" else {"dropout": 0.0}
        model = enc.build_two_tower(kind, embedding, hidden_dim=H, tied_weights=True, **kw)
        model.eval()        # dropout identity; LayerNorm has no running stats
        tower = model.query_tower
        q_ids, d_ids, n_ids = (make_ids(g, B, L, V), make_ids(g, B, L, V, holes=False),
                               make_ids(g, B, L, V, holes=False))
        save = {"q_ids": npy(q_ids), "d_ids": npy(d_ids), "n_ids": npy(n_ids)}
        for k, v in tower_params(tower, kind).items():
            save[f"param_{k}"] = v
        # raw gather (embeddings.py:40) and pooled intermediates
        save["gather_q"] = npy(embedding(q_ids))
        mask = (q_ids > 0).float().unsqueeze(-1)
        save["pooled_q"] = npy((embedding(q_ids) * mask).sum(1) / (mask.sum(1) + 1e-9))
        for loss_name in ["in_batch", "triplet", "multiple_negatives"]:
            model.zero_grad(set_to_none=True)
            qv, dv, nv = model(q_ids, d_ids, n_ids)
            for t in (qv, dv, nv):
                t.retain_grad()
            if loss_name == "in_batch":
                loss = los.in_batch_sampled_softmax_loss(qv, dv, temperature=0.1)
            elif loss_name == "triplet":
                loss = los.build("triplet", margin=0.2)(qv, dv, nv)
            else:
                negs = torch.stack([nv, dv.roll(1, 0), dv.roll(2, 0)], dim=1)   # [B,3,H]
                negs.retain_grad()
                loss = los.build("multiple_negatives", temperature=0.1)(qv, dv, negs)
            loss.backward()
            pre = f"{loss_name}_"
            save[pre + "loss"] = npy(loss)
            save[pre + "q_out"], save[pre + "d_out"], save[pre + "n_out"] = npy(qv), npy(dv), npy(nv)
            save[pre + "dq_out"] = npy(qv.grad)
            save[pre + "dd_out"] = npy(dv.grad)
            if nv.grad is not None:
                save[pre + "dn_out"] = npy(nv.grad)
            if loss_name == "multiple_negatives":
                save[pre + "dnegs"] = npy(negs.grad)
            for k, v in grads_of(tower, kind).items():
                save[pre + f"grad_{k}"] = v
        np.savez_compressed(f"{OUT}/tower_{name}.npz", **save)

    # ---------------------------------------------------------------- losses on raw (non-unit) rows
    g = torch.Generator().manual_seed(77)
    B, H, N = 12, 24, 4
    q = torch.randn(B, H, generator=g).requires_grad_()
    p = torch.randn(B, H, generator=g).requires_grad_()
    n = torch.randn(B, H, generator=g).requires_grad_()
    negs = torch.randn(B, N, H, generator=g).requires_grad_()
    save = {"q": npy(q), "p": npy(p), "n": npy(n), "negs": npy(negs)}
    for nm, fn in [("in_batch", lambda: los.in_batch_sampled_softmax_loss(q, p, temperature=0.5)),
                   ("triplet", lambda: los.contrastive_triplet_loss(q, p, n, margin=0.3)),
                   ("multiple_negatives", lambda: los.multiple_negatives_loss(q, p, negs, temperature=0.2))]:
        for t in (q, p, n, negs):
            t.grad = None
        loss = fn()
        loss.backward()
        save[f"{nm}_loss"] = npy(loss)
        for tn, t in [("q", q), ("p", p), ("n", n), ("negs", negs)]:
            if t.grad is not None:
                save[f"{nm}_d{tn}"] = npy(t.grad)
    np.savez_compressed(f"{OUT}/losses_raw.npz", **save)

    # ---------------------------------------------------------------- TwoTowerSearch
    docs = ["rockets burn fuel to produce thrust", "the cat sat on the mat",
            "two tower models encode queries and documents", "rockets burn fuel to produce thrust",
            "fuel prices rose again this week", "a cat and a dog", "thrust vectoring on rockets",
            "", "documents and queries share one embedding table", "mat"]
    queries = ["how do rockets work", "cat", "two tower"]
    torch.manual_seed(3)
    ctok = tok.build("char").fit(docs + queries)
    embedding = emb.build("lookup", ctok.vocab_size, embedding_dim=16)
    model = enc.build_two_tower("mean", embedding, hidden_dim=32, tied_weights=False)
    s = TwoTowerSearch(model, ctok, device="cpu")
    s.index_documents(docs)
    sres = {"docs": docs, "queries": queries, "vocab": ctok.string_to_index, "results": {}}
    for qq in queries:
        for k in (1, 3, 5, 50):
            r = s.search(qq, top_k=k)
            sres["results"][f"{qq}|{k}"] = [[x["document"], x["score"]] for x in r]
    with open(f"{OUT}/search_small.json", "w") as f:
        json.dump(sres, f, indent=1, sort_keys=True)
    sd = {k: npy(v) for k, v in model.state_dict().items()}
    np.savez_compressed(f"{OUT}/search_small.npz", doc_embeddings=npy(s.document_embeddings),
                        **{k.replace(".", "__"): v for k, v in sd.items()})

    # scoring + topk on a random matrix, with planted exact ties
    g = torch.Generator().manual_seed(7)
    D = torch.nn.functional.normalize(torch.randn(3000, 64, generator=g), dim=-1)
    D[17] = D[5]; D[400] = D[5]; D[2999] = D[5]; D[1000] = D[999]
    Q = torch.nn.functional.normalize(torch.randn(4, 64, generator=g), dim=-1)
    Q[0] = D[5]
    scores = torch.stack([torch.nn.functional.cosine_similarity(
        Q[i:i + 1].unsqueeze(1), D.unsqueeze(0), dim=2).squeeze(0) for i in range(4)])
    tv, ti = torch.topk(scores, 100, dim=1)
    sv, si = torch.sort(scores, dim=1, descending=True, stable=True)
    np.savez_compressed(f"{OUT}/search_topk.npz", D=npy(D), Q=npy(Q), scores=npy(scores),
                        topk_values=npy(tv), topk_indices=npy(ti),
                        stable_values=npy(sv[:, :100]), stable_indices=npy(si[:, :100]))

    # ---------------------------------------------------------------- AdamW + 3 training steps
    torch.manual_seed(5)
    g = torch.Generator().manual_seed(99)
    V, E, H, B, L = 30, 8, 16, 6, 9
    embedding = emb.build("lookup", V, embedding_dim=E)
    model = enc.build_two_tower("mean", embedding, hidden_dim=H, tied_weights=True)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)          # train.py:359
    loss_fn = los.build("triplet", margin=0.2)
    save = {f"init_{k}": v for k, v in tower_params(model.query_tower, "mean").items()}
    for step in range(3):
        qi, di, ni = (make_ids(g, B, L, V), make_ids(g, B, L, V, holes=False),
                      make_ids(g, B, L, V, holes=False))
        qv, dv, nv = model(qi, di, ni)
        loss = loss_fn(qv, dv, nv)
        opt.zero_grad(); loss.backward(); opt.step()               # train.py:137-139
        save[f"step{step}_q_ids"], save[f"step{step}_d_ids"], save[f"step{step}_n_ids"] = npy(qi), npy(di), npy(ni)
        save[f"step{step}_loss"] = npy(loss)
        for k, v in tower_params(model.query_tower, "mean").items():
            save[f"step{step}_{k}"] = v
    np.savez_compressed(f"{OUT}/train_triplet_3steps.npz", **save)

    # same with the in-batch loss (direct call) and untied towers
    torch.manual_seed(6)
    g = torch.Generator().manual_seed(100)
    embedding = emb.build("lookup", V, embedding_dim=E)
    model = enc.build_two_tower("mean", embedding, hidden_dim=H, tied_weights=False)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    save = {f"init_q_{k}": v for k, v in tower_params(model.query_tower, "mean").items()}
    save.update({f"init_d_{k}": v for k, v in tower_params(model.document_tower, "mean").items()})
    for step in range(3):
        qi, di = make_ids(g, B, L, V), make_ids(g, B, L, V, holes=False)
        qv, dv = model(qi, di)
        loss = los.in_batch_sampled_softmax_loss(qv, dv, temperature=0.1)
        opt.zero_grad(); loss.backward(); opt.step()
        save[f"step{step}_q_ids"], save[f"step{step}_d_ids"] = npy(qi), npy(di)
        save[f"step{step}_loss"] = npy(loss)
        for k, v in tower_params(model.query_tower, "mean").items():
            save[f"step{step}_q_{k}"] = v
        for k, v in tower_params(model.document_tower, "mean").items():
            save[f"step{step}_d_{k}"] = v
    np.savez_compressed(f"{OUT}/train_inbatch_untied_3steps.npz", **save)
    print("golden fixtures written to", OUT)
    for fn in sorted(os.listdir(OUT)):
        print(f"  {fn:40s} {os.path.getsize(os.path.join(OUT, fn)):8d} B")


if __name__ == "__main__":
    main()
