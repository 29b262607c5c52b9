"""2-GPU NCCL tests (skipped unless >= 2 CUDA devices): the real kernels under the real collectives.

Equivalence targets: the 2-rank data-parallel step with global in-batch negatives equals the
single-process oracle on the concatenated batch; the row-sharded search equals the unsharded one.
"""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gather_np(t, world):
    import torch.distributed as dist
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t.contiguous())
    return torch.cat(parts, 0).cpu().numpy()


def _trainer_vs_oracle(rank, world, dev, log):
    """ONE data-parallel FusedTrainer step on `world` ranks == the single-process reference step on
    the concatenated batch (twotower/train.py:120-139; losses.py:107-116 in-batch, :28-35 triplet): loss and every
    parameter gradient, for the paths bench.py runs (bf16, global negatives, peer-memory exchange and NCCL) and for the
    triplet / local-negatives data-parallel steps."""
    import torch.distributed as dist
    import two_towers_b200 as tt
    from _parity import BF16_RTOL, check, oracle_step, tower_grads, tower_params, trainer_gates
    B, L, V, E, H = 256, 64, 128, 64, 256
    gq = torch.Generator().manual_seed(100 + rank)
    ids = [torch.randint(0, V, (B, L), generator=gq) for _ in range(3)]
    all_ids = [_gather_np(t.to(dev), world) for t in ids]
    cases = [("in_batch", True, "bf16", True, True), ("in_batch", True, "bf16", False, True), ("in_batch", True, "bf16", True, False),
             ("in_batch", False, "bf16", True, True), ("triplet", False, "bf16", True, True), ("triplet", False, "bf16", True, False),
             ("in_batch", True, "fp32", True, True), ("triplet", False, "fp32", False, True)]
    for loss, glob, prec, p2p, tied in cases:
        torch.manual_seed(0)
        emb = tt.embeddings.build("lookup", V, embedding_dim=E)
        model = tt.build_two_tower("mean", emb, hidden_dim=H, tied_weights=tied).to(dev)
        pq0 = tower_params(model.query_tower)
        pd0 = pq0 if tied else tower_params(model.document_tower)
        tr = tt.FusedTrainer(model, loss=loss, temperature=0.1, margin=2.5, batch_size=B, max_len=L, precision=prec,
                             process_group=dist.group.WORLD, global_negatives=glob, p2p=p2p)
        if loss == "in_batch" and glob and prec == "bf16":
            assert tr.global_fast and tr.p2p == p2p           # the path bench.py times at N > 1
        local = tr.step(*ids[:tr.passes]).clone()
        total = local.clone()
        dist.all_reduce(total)
        gates = None
        if prec == "bf16":
            gates = tuple(_gather_np(torch.tensor(g, device=dev).to(torch.uint8), world).astype(bool) for g in trainer_gates(tr))
        groups = None if (loss != "in_batch" or glob) else [(r * B, (r + 1) * B) for r in range(world)]
        rl, rgq, rgd, flips = oracle_step(pq0, pd0, all_ids[0], all_ids[1], all_ids[2] if loss == "triplet" else None, loss=loss,
                                          temperature=0.1, margin=2.5, gates=gates, groups=groups)
        # in-batch: every rank's loss kernel is scaled by 1 / (world * B_local), so the rank values ADD UP to the reference
        # loss (global negatives) or to the mean of the per-rank losses (local negatives); the triplet kernel reports the
        # local row mean, whose average over ranks is the reference value on the concatenated batch
        got_loss = total.item() / world if loss == "triplet" else total.item()
        tol = BF16_RTOL if prec == "bf16" else 1e-4
        head = (f"  {world}-rank FusedTrainer loss={loss} global_negatives={glob} {prec} p2p={tr.p2p} tied={tied}: "
                f"loss {got_loss:.6f} oracle {rl:.6f} gates flipped {flips:.2%}")
        print(head); log.append(head)
        assert abs(got_loss - rl) <= tol * abs(rl), (got_loss, rl)
        assert flips < 0.01
        got_q = tower_grads(model.query_tower)
        for k in rgq:
            check(got_q[k], rgq[k], tol, f"grad query/{k}", log)
        if not tied:
            got_d = tower_grads(model.document_tower)
            for k in ("w1", "b1", "w2", "b2"):
                # in-batch, untied: db2 of the document tower is a cancellation residue (sum_j dL/dd_j = 0), see
                # tests/test_gpu_tensor_core.py::test_fused_trainer_bf16_tracks_fp32
                kt = 5e-2 if (k == "b2" and loss == "in_batch" and prec == "bf16") else tol
                check(got_d[k], rgd[k], kt, f"grad document/{k}", log)
        both = _gather_np(tr.flat.double().sum().reshape(1), world)
        assert both[0] == both[1]                             # identical parameters on every rank after the step
        del tr, model
    # ---- sharded index build -- concat of the shards == the single-GPU matrix, bit for bit ----------
    torch.manual_seed(0)
    tok = tt.CharTokeniser().fit(["abcdefghijklmnopqrstuvwxyz 0123456789"])
    emb = tt.embeddings.build("lookup", tok.vocab_size, embedding_dim=E)
    model = tt.build_two_tower("mean", emb, hidden_dim=H, tied_weights=True).to(dev)
    docs = [f"document number {i} about topic {i % 17} and {(i * 7) % 13}" for i in range(1001)]
    for idt in ("fp32", "bf16"):
        s = tt.TwoTowerSearch(model, tok, device=dev, index_dtype=idt, process_group=dist.group.WORLD, encode_batch_size=128)
        s.index_documents(docs)
        lo, hi = tt.parallel.shard_bounds(len(docs), rank, world)
        assert s.row_offset == lo and tuple(s.document_embeddings.shape) == (hi - lo, H) and s.num_documents == len(docs)
        full = s._encode(docs, model.document_tower)          # what a single GPU builds (same tower, same kernels)
        if idt == "bf16":
            full = tt.ops.cast_bf16(full)
        assert torch.equal(s.document_embeddings, full[lo:hi])
        res = s.search("document number 500 about topic 7 and 3", top_k=5)
        one = tt.ops.topk_scan(full, s._encode(["document number 500 about topic 7 and 3"], model.query_tower), 5, cosine=True)
        assert [r["document"] for r in res] == [docs[i] for i in one[1][0].tolist()]
    log.append("  sharded index_documents: shard rows == single-GPU matrix rows (fp32 and bf16), sharded search == unsharded")


def _worker(rank, world, port, q_out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import ctypes as C
        import two_towers_b200 as tt
        from oracle import two_tower_oracle as O
        from two_towers_b200 import parallel, _lib
        # ---- global in-batch negatives: loss / dq / dd vs the oracle on the concatenated batch -------------
        g = np.random.default_rng(0)
        B, H, t = 192, 256, 0.1
        Q = O.normalize(g.standard_normal((world * B, H))).astype(np.float32)
        D = O.normalize(g.standard_normal((world * B, H))).astype(np.float32)
        ql = torch.tensor(Q[rank * B:(rank + 1) * B], device=dev)
        dl = torch.tensor(D[rank * B:(rank + 1) * B], device=dev)
        for prec, tol in (("fp32", 2e-5), ("bf16", 2e-2)):
            loss, lse, dglob = parallel.global_inbatch_fwd(ql, dl, t, tt.ops, precision=prec)
            total = loss.clone()
            dist.all_reduce(total)
            rl, rlse = O.in_batch_loss(Q.astype(np.float64), D.astype(np.float64), t)
            assert abs(total.item() - rl) <= tol * abs(rl), (prec, total.item(), rl)
            dq, dd = parallel.global_inbatch_bwd(ql, dl, dglob, lse, t, tt.ops, precision=prec)
            rdq, rdd = O.in_batch_loss_bwd(Q.astype(np.float64), D.astype(np.float64), t)
            for a, b in ((dq, rdq), (dd, rdd)):
                b = b[rank * B:(rank + 1) * B]
                err = np.abs(a.cpu().numpy() - b).max()
                assert err <= tol * np.abs(b).max() + 1e-9, (prec, err, np.abs(b).max())
        # ---- fused trainer, 2 ranks, global negatives: finite, decreasing, identical weights on both ranks --
        torch.manual_seed(0)
        emb = tt.embeddings.build("lookup", 128, embedding_dim=64)
        model = tt.build_two_tower("mean", emb, hidden_dim=256, tied_weights=True).to(dev)
        gq = torch.Generator().manual_seed(100 + rank)
        q_ids = torch.randint(1, 128, (256, 64), generator=gq)
        d_ids = torch.randint(1, 128, (256, 64), generator=gq)
        tr = tt.FusedTrainer(model, loss="in_batch", batch_size=256, max_len=64, precision="bf16",
                             process_group=dist.group.WORLD, global_negatives=True)
        first = tr.step(q_ids, d_ids).item()
        for _ in range(10):
            last = tr.step(q_ids, d_ids).item()
        assert np.isfinite(first) and np.isfinite(last) and last < first
        chk = tr.flat.double().sum().reshape(1)
        both = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(both, chk)
        assert both[0].item() == both[1].item()              # ranks stay bit-identical after the all-reduce
        # ---- NVLink peer-memory exchange (tt_p2p_allgather): raw all-gather, repeated rounds, rank-order sum -----------
        x = parallel.P2PExchange(4096 * 4, device=dev)
        for rnd in range(3):
            src = torch.full((4096,), float(10 * rnd + rank + 1), device=dev)
            x.allgather(src)
            got = x.gathered(torch.float32, (4096,))
            torch.cuda.synchronize()
            for r in range(world):
                assert float(got[r].min()) == float(got[r].max()) == 10 * rnd + r + 1, (rnd, r, got[r][:4])
            out = torch.empty(4096, device=dev)
            x.sum_slots(out)
            assert float(out[0]) == sum(10 * rnd + r + 1 for r in range(world))
            dist.barrier()                                   # single-buffered: nobody pushes the next round while a peer still reads
        x2 = parallel.P2PExchange(1024 * 4, device=dev, double_buffered=True)
        for rnd in range(5):                                 # double-buffered: back-to-back rounds need no barrier in between
            x2.allgather(torch.full((1024,), float(rnd * 100 + rank), device=dev))
            out = torch.empty(1024, device=dev)
            x2.sum_slots(out)
            assert float(out[7]) == sum(rnd * 100 + r for r in range(world)), (rnd, float(out[7]))
        # ---- the same trainer with NCCL collectives and with peer-memory exchanges: identical trajectories -------------
        def run(p2p):
            torch.manual_seed(0)
            emb2 = tt.embeddings.build("lookup", 128, embedding_dim=64)
            m = tt.build_two_tower("mean", emb2, hidden_dim=256, tied_weights=True).to(dev)
            t2 = tt.FusedTrainer(m, loss="in_batch", batch_size=256, max_len=64, precision="bf16",
                                 process_group=dist.group.WORLD, global_negatives=True, p2p=p2p)
            assert t2.p2p == p2p
            losses = [t2.step(q_ids, d_ids).item() for _ in range(6)]
            return losses, t2.flat.clone()
        l_nccl, w_nccl = run(False)
        l_p2p, w_p2p = run(True)
        assert l_nccl == l_p2p, (l_nccl, l_p2p)               # 2 ranks: a + b in either order is the same float
        assert torch.equal(w_nccl, w_p2p)
        # ---- sharded search == unsharded ---------------------------------------------------------------------
        N, k = 100_003, 50
        idx = O.normalize(g.standard_normal((N, H))).astype(np.float32)
        idx[70_000] = idx[3]
        qs = O.normalize(g.standard_normal((3, H))).astype(np.float32); qs[0] = idx[3]
        lo, hi = parallel.shard_bounds(N, rank, world)
        s, i = parallel.sharded_topk(torch.tensor(idx[lo:hi], device=dev), torch.tensor(qs, device=dev), k, lo, tt.ops, cosine=False)
        fs, fi = tt.ops.topk_scan(torch.tensor(idx, device=dev), torch.tensor(qs, device=dev), k, cosine=False)
        assert torch.equal(i, fi) and torch.equal(s, fs)
        assert i[0, :2].tolist() == [3, 70_000]
        # graph-replayed sharded search: same answers on every call (both buffer parities, fresh queries)
        st = parallel.ShardedTopK(torch.tensor(idx[lo:hi], device=dev), k, lo, tt.ops, cosine=False, nq=3)
        for rep in range(5):
            qr = torch.tensor(np.roll(qs, rep, axis=0), device=dev)
            s2, i2 = st(qr)
            fs2, fi2 = tt.ops.topk_scan(torch.tensor(idx, device=dev), qr, k, cosine=False)
            assert torch.equal(i2, fi2) and torch.equal(s2, fs2), rep
        assert st.fused                                      # scan + ONE merge / exchange / merge kernel (tt_topk_scan_p2p)
        # one query, k = 100 (the benchmarked form), more rounds than buffer parities; then the three-launch chain it replaces
        st1 = parallel.ShardedTopK(torch.tensor(idx[lo:hi], device=dev), 100, lo, tt.ops, cosine=True, nq=1)
        assert st1.fused
        os.environ["TT_SEARCH_FUSED"] = "0"
        st0 = parallel.ShardedTopK(torch.tensor(idx[lo:hi], device=dev), 100, lo, tt.ops, cosine=True, nq=1)
        os.environ.pop("TT_SEARCH_FUSED")
        assert not st0.fused
        for rep in range(9):
            qr = torch.tensor(qs[rep % 3:rep % 3 + 1] if rep < 6 else idx[70_000 + rep:70_001 + rep], device=dev)
            fs1, fi1 = tt.ops.topk_scan(torch.tensor(idx, device=dev), qr, 100, cosine=True)
            s1, i1 = st1(qr)
            assert torch.equal(i1, fi1) and torch.equal(s1, fs1), rep
            s0, i0 = st0(qr)
            assert torch.equal(i0, fi1) and torch.equal(s0, fs1), rep
        rc = C.c_int(-1)
        _lib.check(_lib.load().tt_p2p_status(C.byref(st1.xch.desc), C.byref(rc)), "tt_p2p_status")
        assert rc.value == -1
        log = []
        _trainer_vs_oracle(rank, world, dev, log)
        if rank == 0:
            out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
            os.makedirs(out_dir, exist_ok=True)
            with open(os.path.join(out_dir, "multi_parity.log"), "w") as f:
                f.write("\n".join(log) + "\n")
        q_out.put((rank, "ok"))
    except Exception:
        import traceback
        q_out.put((rank, traceback.format_exc()))
    finally:
        # results are already in the queue: leave without the (occasionally slow) NCCL / CUDA-graph teardown
        torch.cuda.synchronize()
        q_out.close(); q_out.join_thread()
        os._exit(0)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_nccl_equivalence():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(30)
        if p.is_alive():
            p.kill()
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}:\n{msg}"
