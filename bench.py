#!/usr/bin/env python
"""bench.py -- the BASELINE.json metric on B200: two-tower train pairs/sec + top-k search QPS.

    python bench.py --gpus N --steps K --warmup W [--impl ours|reference] [--workload train|search]
                    [--precision bf16|fp32]

One rank per GPU (torchrun for N>1, env RANK/LOCAL_RANK/WORLD_SIZE/MASTER_*).  Rank 0 prints ONE
JSON line.  The headline `metric` is train query-doc pairs/sec on configs[1]
(configs/char_tower.yml shape: char-level tied mean towers, in-batch softmax, B=4096/GPU, L=64,
E=64, d=256, AdamW); the same line carries a `search` object for the second half of the
BASELINE metric (top-100 QPS over a synthetic 10M x 256 index, row-sharded over the ranks).

  value     device-timed whole-job pairs/s: inputs resident in HBM, CUDA events around each step
            on the launching stream, L2 flushed between timed steps, max over ranks.
  e2e       same metric through the public API (FusedTrainer.step) with pinned HOST id tensors:
            H2D of the ids and D2H of the loss inside the timed region.
  roofline  dominant kernel (the fused in-batch CE backward): algorithmic FLOPs / live
            CUDA-event time of that kernel vs MEASURED_PEAKS.json.
  cpu_baseline  oracle/torch_port.py (the reference's eager path restated) on the host cores.
`--impl reference` times that same CPU port as the reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

# configs[1]: configs/char_tower.yml (tokeniser.max_len 64, embedding_dim 64, tied mean towers),
# north_star fixes B=4096, d=256, in-batch softmax; synthetic char vocabulary V=128.
CFG = dict(V=128, L=64, E=64, H=256, B=4096, temperature=0.1, lr=1e-3)
SEARCH = dict(N=10_000_000, H=256, k=100)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for nm, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def synth_ids(B, L, V, seed):
    """lengths ~U{8..L}, ids ~U{1..V-1}, zero padded."""
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(8, L + 1, (B, 1), generator=g)
    ids = torch.randint(1, V, (B, L), generator=g)
    # int32 on the wire: tokenisers.encode_batch(out=<pinned int32 buffer>) fills such arrays directly and the kernels read
    # int32 ids as they are (the reference's int64 ids are a torch default, not a requirement -- SURVEY 8f-2)
    return torch.where(torch.arange(L)[None, :] < lens, ids, torch.zeros_like(ids)).to(torch.int32)


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the eager path restated in oracle/torch_port.py, on host cores
# ------------------------------------------------------------------------------------------
def cpu_train_baseline(steps, warmup):
    from oracle import torch_port as P
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = P.PortTwoTower(CFG["V"], CFG["E"], CFG["H"], tied=True)
    opt = torch.optim.AdamW(model.parameters(), lr=CFG["lr"])
    q, d = synth_ids(CFG["B"], CFG["L"], CFG["V"], 1234).long(), synth_ids(CFG["B"], CFG["L"], CFG["V"], 4321).long()   # torch.long, as the reference feeds nn.Embedding
    for _ in range(warmup):
        P.train_step(model, opt, "in_batch", q, d, temperature=CFG["temperature"])
    t0 = time.perf_counter()
    for _ in range(steps):
        P.train_step(model, opt, "in_batch", q, d, temperature=CFG["temperature"])
    dt = (time.perf_counter() - t0) / steps
    return dict(value=CFG["B"] / dt, unit="pairs/s", cores=cores, kind="port", ms_per_step=dt * 1e3,
                sample=f"{steps} full train steps (B={CFG['B']}, same config) of oracle/torch_port.py, torch CPU "
                       f"{torch.__version__}, {cores} threads")


def cpu_search_baseline(reps=3, n=1_000_000):
    from oracle import torch_port as P
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(7)
    D = torch.nn.functional.normalize(torch.randn(n, SEARCH["H"], generator=g), dim=-1)
    q = torch.nn.functional.normalize(torch.randn(1, SEARCH["H"], generator=g), dim=-1)
    P.search(q, D, SEARCH["k"])
    t0 = time.perf_counter()
    for _ in range(reps):
        P.search(q, D, SEARCH["k"])
    dt = (time.perf_counter() - t0) / reps
    return dict(value=1.0 / dt, unit="queries/s", cores=cores, kind="port",
                sample=f"{reps} queries, reference cosine+topk formulation over N={n} rows (1/10 of the 10M index); "
                       f"extrapolated 10M-row QPS = {1.0 / dt / (SEARCH['N'] / n):.3f}",
                qps_at_10M_extrapolated=1.0 / dt / (SEARCH["N"] / n))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 30)), max(1, min(args.warmup, 3))
    base = cpu_train_baseline(steps, warmup)
    line = {
        "impl": "reference", "metric": "train query-doc pairs/sec", "value": base["value"], "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": base["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.gpus, "fp32"),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.workload == "search":
        sb = cpu_search_baseline()
        line.update(metric="top-k search QPS over 10M-doc index", value=sb["qps_at_10M_extrapolated"],
                    unit="queries/s", cpu_baseline={k: sb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                    e2e={"value": sb["qps_at_10M_extrapolated"], "unit": "queries/s", "h2d_bytes_per_step": 0,
                         "d2h_bytes_per_step": 0})
    print(json.dumps(line), flush=True)


def config_dict(n_gpus, precision):
    return {"workload": "configs/char_tower.yml shape: tied mean towers, in-batch softmax (tau=0.1), AdamW(1e-3); "
                        f"V={CFG['V']} L={CFG['L']} E={CFG['E']} d={CFG['H']} B={CFG['B']}/GPU",
            "global_batch": CFG["B"] * n_gpus, "seq_len": CFG["L"], "parallelism": f"dp{n_gpus}",
            "negatives": "global in-batch (all-gather D)" if n_gpus > 1 else "in-batch",
            "precision_mode": precision, "l2": "flushed (256 MiB write) between timed steps"}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def flush_l2(buf):
    buf.add_(1)


def bench_train(args, dev, rank, world, pg):
    import two_towers_b200 as tt
    torch.manual_seed(0)
    emb = tt.embeddings.build("lookup", CFG["V"], embedding_dim=CFG["E"])
    model = tt.build_two_tower("mean", emb, hidden_dim=CFG["H"], tied_weights=True).to(dev)
    tr = tt.FusedTrainer(model, loss="in_batch", temperature=CFG["temperature"], lr=CFG["lr"], batch_size=CFG["B"],
                         max_len=CFG["L"], precision=args.precision, process_group=pg, global_negatives=True,
                         id_dtype=torch.int32)
    B, L, V = CFG["B"], CFG["L"], CFG["V"]
    host_q = [synth_ids(B, L, V, 1234 + rank + 100 * i).pin_memory() for i in range(4)]
    host_d = [synth_ids(B, L, V, 4321 + rank + 100 * i).pin_memory() for i in range(4)]
    dev_q = [t.to(dev) for t in host_q]
    dev_d = [t.to(dev) for t in host_d]
    launches_per_step = None
    if rank == 0 or True:
        tr.load_batch(dev_q[0], dev_d[0])
        launches_per_step = tr.kernels_per_step()
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev)       # 256 MiB > 126 MB L2
    for i in range(max(args.warmup, 3)):
        tr.load_batch(dev_q[i % 4], dev_d[i % 4])
        tr.run()
    torch.cuda.synchronize()
    # ---- device-timed: inputs resident, events around each step, L2 flushed between ----------
    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    evs = []
    for i in range(args.steps):
        tr.load_batch(dev_q[i % 4], dev_d[i % 4])
        flush_l2(flush)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tr.run()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    t_dev = float(np.sum(step_ms)) / 1e3
    # ---- end to end: pinned host ids -> H2D -> step -> D2H loss, wall clock over K steps ----------
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    t0 = time.perf_counter()
    last = 0.0
    tr.prefetch(host_q[0], host_d[0])                       # every step's H2D copy is inside the timed region;
    pending = None
    for i in range(args.steps):                             # the copy of batch i+1 overlaps the compute of batch i
        tr.step()
        nxt = tr.read_loss_async()                          # D2H read of EVERY step's loss (pinned ring), consumed one
        if i + 1 < args.steps:                              # step late so the host never stalls the GPU
            tr.prefetch(host_q[(i + 1) % 4], host_d[(i + 1) % 4])
        if pending is not None:
            last = pending()
        pending = nxt
    last = pending()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    # ---- dominant kernel timed alone on the launching stream (live, CUDA events) ------------
    lib = tt._lib.load()
    y = tr.y
    roof = kernel_roofline(tt, tr, dev)
    t = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    t_dev, t_e2e = t.tolist()
    return dict(t_dev=t_dev, t_e2e=t_e2e, clocks=clocks, launches_per_step=launches_per_step, roof=roof,
                loss=last, h2d=2 * B * L * 4, d2h=4)


def time_train_local_negatives(args, dev, rank, world, pg):
    """Secondary multi-GPU number: the same step with LOCAL in-batch negatives (each rank's loss over its own batch,
    gradients averaged by allreduce -- what wrapping the reference loop in DDP gives).  Per-GPU work is then independent
    of the world size, unlike global negatives whose loss FLOPs grow with it."""
    import two_towers_b200 as tt
    torch.manual_seed(0)
    emb = tt.embeddings.build("lookup", CFG["V"], embedding_dim=CFG["E"])
    model = tt.build_two_tower("mean", emb, hidden_dim=CFG["H"], tied_weights=True).to(dev)
    tr = tt.FusedTrainer(model, loss="in_batch", temperature=CFG["temperature"], lr=CFG["lr"], batch_size=CFG["B"],
                         max_len=CFG["L"], precision=args.precision, process_group=pg, global_negatives=False,
                         id_dtype=torch.int32)
    B, L, V = CFG["B"], CFG["L"], CFG["V"]
    dev_q = [synth_ids(B, L, V, 1234 + rank + 100 * i).to(dev) for i in range(4)]
    dev_d = [synth_ids(B, L, V, 4321 + rank + 100 * i).to(dev) for i in range(4)]
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    for i in range(max(args.warmup, 3)):
        tr.load_batch(dev_q[i % 4], dev_d[i % 4])
        tr.run()
    torch.cuda.synchronize()
    torch.distributed.barrier()
    torch.cuda.synchronize()
    evs = []
    for i in range(args.steps):
        tr.load_batch(dev_q[i % 4], dev_d[i % 4])
        flush_l2(flush)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); tr.run(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    torch.distributed.barrier()
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs) / 1e3], dtype=torch.float64, device=dev)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def kernel_roofline(tt, tr, dev):
    """Time the loss backward (dominant kernel of the step) alone, L2-cold, with CUDA events: the same single launch the
    trainer issues (both gradient passes, fused with the normalise backward when the shape allows)."""
    import ctypes as C
    from two_towers_b200 import _lib
    pk = peaks()
    lib = _lib.load()
    Bl, H = tr.B, tr.H
    Bg = tr.B * tr.world
    prec = "bf16" if tr.prec == 1 else "fp32"
    q = torch.nn.functional.normalize(torch.randn(Bl, H, device=dev), dim=-1)
    d = torch.nn.functional.normalize(torch.randn(Bg, H, device=dev), dim=-1)
    loss, lse, _ = tt.ops.inbatch_ce_fwd(q, d, 0.1, precision=prec)
    merged = prec == "bf16" and H % 64 == 0 and H <= 256
    if merged:
        # rank-local view of the data-parallel step: local queries / local documents against all Bg rows
        qb, db = tt.ops.cast_bf16(q), tt.ops.cast_bf16(d)
        lse_g = torch.cat([lse] * tr.world) if tr.world > 1 else lse
        n = int(lib.tt_inbatch_ce_bwd_nparts_ex(Bl, Bg, Bl, Bg, H))
        fused = bool(lib.tt_inbatch_ce_bwd_fused_ok(Bl, Bg, Bl, Bg, H)) and Bl % 32 == 0
        vp = lambda t: None if t is None else t.data_ptr()
        inv = torch.ones(Bl, device=dev)
        if fused:
            dzq = torch.empty(Bl, H, dtype=torch.bfloat16, device=dev); dzd = torch.empty_like(dzq)
            csq = torch.empty(Bl // 32, H, device=dev); csd = torch.empty_like(csq)
            qp = _lib.CePass(vp(qb), Bl, vp(db), Bg, Bg, Bg, 0, 0, vp(lse), 0, None, 0, vp(dzq), vp(csq), vp(inv))
            dp = _lib.CePass(vp(db), Bl, vp(qb if tr.world == 1 else db), Bg, Bg, Bg, 0, 0, vp(lse_g), 0, None, 0, vp(dzd), vp(csd), vp(inv))
        else:
            pq_ = torch.empty(n, Bl, H, device=dev); pd_ = torch.empty(n, Bl, H, device=dev)
            qp = _lib.CePass(vp(qb), Bl, vp(db), Bg, Bg, Bg, 0, 0, vp(lse), 0, vp(pq_), Bl * H, None, None, None)
            dp = _lib.CePass(vp(db), Bl, vp(qb if tr.world == 1 else db), Bg, Bg, Bg, 0, 0, vp(lse_g), 0, vp(pd_), Bl * H, None, None, None)
        def run():
            s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(lib.tt_inbatch_ce_bwd_parts_ex(C.byref(qp), C.byref(dp), H, 10.0, 1.0 / Bg, None, n, s), "ce_bwd")
        what = f"inbatch_ce_bwd[{prec}] (dQ+dD in one launch, fused recompute" + (", fused normalise backward)" if fused else ")")
    else:
        def run():
            tt.ops.inbatch_ce_bwd(q, d, lse, 0.1, precision=prec)
        what = f"inbatch_ce_bwd[{prec}] (dQ+dD, fused recompute)"
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    # capture the op in a CUDA graph so the CUDA-event interval holds device time only (no host launch gaps)
    run()
    torch.cuda.synchronize()
    NREP = 10                                   # launches per timed replay: amortises the ~8-10 us replay floor of a
    graph = torch.cuda.CUDAGraph()              # one-kernel graph, which would otherwise be charged to the kernel
    with torch.cuda.graph(graph):
        for _ in range(NREP):
            run()
    times = []
    for i in range(16):
        flush_l2(flush)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        if i >= 4:
            times.append(e0.elapsed_time(e1) / NREP)
    ms = float(np.mean(times))
    flops = 4.0 * Bl * Bg * H                   # algorithmic backward FLOPs (dQ + dD products); recompute not counted
    achieved = flops / (ms * 1e-3) / 1e12
    # DRAM bytes of this launch from the committed ncu --set full capture (profiles/r01_ncu_ce_bwd_final_raw.csv:
    # dram__bytes_read.sum 4.42 MB + dram__bytes_write.sum 0 -- the 8 MB of outputs stay in the 126 MB L2); only that
    # shape was captured
    traffic = 4416512 if (merged and Bl == 4096 and Bg == 4096 and H == 256) else None
    return {"kernel": what, "bound": "tensor", "achieved": achieved,
            "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": achieved / pk["tf_burst"], "traffic": traffic,
            "ms": ms, "launch_flops": flops, "executed_flops": 2 * flops,
            "note": "achieved counts algorithmic FLOPs; the kernel executes 2x (S = X Y^T is recomputed, flash style); "
                    f"duration = CUDA-event time of {NREP} back-to-back launches in one graph replay / {NREP} (L2 flushed before each replay)",
            "peak_source": f"{pk['src']} bf16 burst (kernel timed alone)"}


def bench_search(args, dev, rank, world, pg):
    import two_towers_b200 as tt
    from two_towers_b200 import parallel
    N, H, k = SEARCH["N"], SEARCH["H"], SEARCH["k"]
    lo, hi = parallel.shard_bounds(N, rank, world)
    gen = torch.Generator(device=dev).manual_seed(7 + rank)
    out = {}
    for idx_dtype in ("fp32", "bf16"):
        D = torch.empty(hi - lo, H, device=dev, dtype=torch.float32)
        for a in range(0, hi - lo, 1_000_000):
            b = min(a + 1_000_000, hi - lo)
            D[a:b] = torch.nn.functional.normalize(torch.randn(b - a, H, device=dev, generator=gen), dim=-1)
        index = D if idx_dtype == "fp32" else tt.ops.cast_bf16(D)
        del D
        qs = torch.nn.functional.normalize(torch.randn(64, H, device=dev, generator=torch.Generator(device=dev).manual_seed(11)), dim=-1)
        ws = torch.empty(tt.ops.topk_scan_workspace_bytes(hi - lo, H, 1, k), dtype=torch.uint8, device=dev)
        nrep = max(10, args.steps)

        sharded = parallel.ShardedTopK(index, k, lo, tt.ops, pg, cosine=False, nq=1) if world > 1 else None

        def one(i):
            q = qs[i % 64:i % 64 + 1]
            if world > 1:                                   # one graph replay: scan, peer-memory candidate exchange, merge
                return sharded(q)
            return tt.ops.topk_scan(index, q, k, cosine=False, id_offset=lo, workspace=ws)
        for i in range(3):
            one(i)
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(nrep):
            one(i)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 1e3 / nrep], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        sec = t.item()
        bytes_per_query = (hi - lo) * H * (4 if idx_dtype == "fp32" else 2)
        pk = peaks()
        ach = bytes_per_query / sec / 1e9
        out[idx_dtype] = {"qps": 1.0 / sec, "ms_per_query": sec * 1e3,
                          "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                                       "frac": ach / pk["hbm"], "traffic": None,
                                       "bytes_per_query_per_gpu": bytes_per_query, "peak_source": pk["src"]}}
        # end to end through the public API: query string -> tokenise -> H2D -> tower -> scan -> D2H -> dicts
        if world == 1 and idx_dtype == "fp32":
            tok = tt.CharTokeniser().fit(["abcdefghijklmnopqrstuvwxyz 0123456789"])
            emb = tt.embeddings.build("lookup", tok.vocab_size, embedding_dim=64)
            model = tt.build_two_tower("mean", emb, hidden_dim=H, tied_weights=True).to(dev)

            class LazyDocs:
                def __len__(self):
                    return N

                def __getitem__(self, i):
                    return f"doc-{i}"
            s = tt.TwoTowerSearch(model, tok, device=dev, cosine=True)
            s.set_index(index, LazyDocs())
            for i in range(3):
                s.search("how do rockets work", top_k=k)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(nrep):
                res = s.search(f"how do rockets work {i}", top_k=k)
            dt = (time.perf_counter() - t0) / nrep
            out["e2e"] = {"value": 1.0 / dt, "unit": "queries/s", "h2d_bytes_per_step": 64 * 8,
                          "d2h_bytes_per_step": k * 12, "api": "TwoTowerSearch.search(str, top_k=100), fp32 index, cosine"}
        del index
        torch.cuda.empty_cache()
    out["config"] = {"N": N, "d": H, "k": k, "shards": world, "queries": "single query per call (reference API)",
                     "scores": "dot product on unit rows", "l2": "index (>= 640 MB/GPU) larger than L2"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "search"])
    ap.add_argument("--precision", default=os.environ.get("TT_BENCH_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--no-search", action="store_true", help="skip the search section of the train line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference_arm(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
        pg = torch.distributed.group.WORLD
    import two_towers_b200 as tt
    try:
        tr = bench_train(args, dev, rank, world, pg)
    except RuntimeError as e:
        if args.precision == "bf16" and "not built" in str(e):
            args.precision = "fp32"
            tr = bench_train(args, dev, rank, world, pg)
        else:
            raise
    t_local = time_train_local_negatives(args, dev, rank, world, pg) if world > 1 else None
    search = None if args.no_search else bench_search(args, dev, rank, world, pg)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_train_baseline(steps=30, warmup=2)
        if search is not None:
            search["cpu_baseline"] = cpu_search_baseline()
    if rank == 0:
        gb = CFG["B"] * world
        K = args.steps
        line = {
            "metric": "train query-doc pairs/sec", "value": gb * K / tr["t_dev"], "unit": "pairs/s",
            "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3), "ms_per_step": tr["t_dev"] / K * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": config_dict(world, args.precision),
            "e2e": {"value": gb * K / tr["t_e2e"], "unit": "pairs/s", "h2d_bytes_per_step": tr["h2d"],
                    "d2h_bytes_per_step": tr["d2h"], "api": "FusedTrainer.prefetch(pinned int32 q_ids, d_ids) / .step() / .read_loss_async() -- H2D of batch i+1 overlaps step i, every loss is read on the host one step late"},
            "gpu_launches": int(tr["launches_per_step"]) * K,
            "gpu_launches_per_step": int(tr["launches_per_step"]),
            "clocks": tr["clocks"], "roofline": tr["roof"], "final_loss": tr["loss"],
            "step_roofline": {"flops_per_step_per_gpu": 2 * 6 * CFG["B"] * (CFG["E"] * CFG["H"] + CFG["H"] ** 2) + 6 * CFG["B"] * gb * CFG["H"],
                              "note": "algorithmic FLOPs (SURVEY 8d) / device step time vs sustained bf16 peak"},
        }
        fl = line["step_roofline"]["flops_per_step_per_gpu"]
        pk = peaks()
        line["step_roofline"]["achieved_tflops"] = fl / (tr["t_dev"] / K) / 1e12
        line["step_roofline"]["frac"] = line["step_roofline"]["achieved_tflops"] / pk["tf_sust"]
        if t_local is not None:
            line["local_negatives"] = {"value": gb * K / t_local, "unit": "pairs/s", "ms_per_step": t_local / K * 1e3,
                                       "note": "same step with per-rank in-batch negatives (DDP semantics): per-GPU work does not grow "
                                               "with the world size; the headline `value` uses GLOBAL negatives, whose loss FLOPs do"}
        if cpu is not None:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if search is not None:
            line["search"] = search
        if args.workload == "search" and search is not None:
            line.update(metric="top-k search QPS over 10M-doc index", value=search["fp32"]["qps"], unit="queries/s",
                        ms_per_step=search["fp32"]["ms_per_query"], roofline=search["fp32"]["roofline"], dtype="f32")
            if "e2e" in search:
                line["e2e"] = search["e2e"]
            if "cpu_baseline" in search:
                line["cpu_baseline"] = {k: search["cpu_baseline"][k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
