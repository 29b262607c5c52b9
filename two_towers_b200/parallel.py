"""Multi-GPU pieces (one process per GPU, ``torch.distributed`` / NCCL over NVLink).

The reference has no distributed code at all; the path shards naturally in two
places and only these use a collective:

* training with global in-batch negatives: all-gather of the document embeddings before the
  fused loss kernel (``label_offset = rank * B_local``), all-gather of (Q, lse) for the dD
  pass, one all-reduce of the flat gradient buffer;
* search: row-sharded index, local top-k, all-gather of k candidates per rank, merge kernel.

The numerical work is delegated to a ``kernels`` namespace (default: ``two_towers_b200.ops``,
i.e. the CUDA library).  The world_size-2 ``gloo`` CPU tests inject a CPU stand-in namespace
to exercise exactly this host logic without a GPU.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def all_gather_rows(x: torch.Tensor, group=None) -> torch.Tensor:
    """[B_local, ...] -> [world * B_local, ...] (rank-major), no autograd."""
    rank, ws = world(group)
    if ws == 1:
        return x
    out = torch.empty((ws * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def global_inbatch_fwd(q_local, d_local, temperature: float, kernels, group=None, precision=None):
    """Forward of the global-negatives in-batch loss on this rank's rows.

    Returns (loss_local_contribution, lse_local, d_global).  The *global* loss is the SUM of
    loss_local_contribution over ranks (each already scaled by 1/B_global), which equals the
    reference ``in_batch_sampled_softmax_loss`` on the concatenated batch.
    """
    rank, ws = world(group)
    B_local = q_local.shape[0]
    d_global = all_gather_rows(d_local, group)
    loss, lse, _ = kernels.inbatch_ce_fwd(q_local, d_global, temperature, label_offset=rank * B_local,
                                          loss_scale=1.0 / (ws * B_local), precision=precision)
    return loss, lse, d_global


def global_inbatch_bwd(q_local, d_local, d_global, lse_local, temperature: float, kernels, group=None,
                       precision=None, grad_out=None):
    """Backward: dq for local queries (vs all docs) and dd for local docs (vs all queries).

    dd uses the all-gather(Q, lse) + local recompute variant: same bytes on the wire as the
    forward, no fp32 reduce-scatter, and deterministic.
    """
    rank, ws = world(group)
    B_local = q_local.shape[0]
    scale = 1.0 / (ws * B_local)
    dq, _ = kernels.inbatch_ce_bwd(q_local, d_global, lse_local, temperature, label_offset=rank * B_local,
                                   loss_scale=scale, grad_out=grad_out, need_dq=True, need_dd=False,
                                   precision=precision)
    q_global = all_gather_rows(q_local, group)
    lse_global = all_gather_rows(lse_local, group)
    _, dd = kernels.inbatch_ce_bwd(q_global, d_local, lse_global, temperature, label_offset=-rank * B_local,
                                   loss_scale=scale, grad_out=grad_out, need_dq=False, need_dd=True,
                                   precision=precision)
    return dq, dd


def allreduce_sum_(flat: torch.Tensor, group=None) -> torch.Tensor:
    _, ws = world(group)
    if ws > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


class _DevMem:
    """Minimal __cuda_array_interface__ carrier so torch can view a raw device pointer."""
    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}


class P2PExchange:
    """All-gather over NVLink peer memory (tt_p2p_allgather): one kernel per exchange, no NCCL on the data path.

    Every rank owns a cudaMalloc'ed buffer of `world` slots, exports it through CUDA IPC and maps every peer's.  The
    handles travel once at construction (torch.distributed all_gather_object); afterwards an exchange is a single
    graph-capturable kernel launch.  `gathered(dtype, shape)` views the local buffer's slots as a tensor.
    """

    def __init__(self, slot_bytes: int, group=None, device=None, double_buffered: bool = False, timeout_s: int = 0):
        from . import _lib
        import ctypes as C
        self.lib = _lib.load()
        self.rank, self.world = world(group)
        if self.world > 8:
            raise RuntimeError("P2PExchange supports up to 8 ranks (one NVSwitch domain)")
        self.group = group
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.slot_bytes = (int(slot_bytes) + 255) // 256 * 256
        self.double_buffered = bool(double_buffered)
        self.nbytes = int(self.lib.tt_p2p_buffer_bytes(self.world, self.slot_bytes, int(self.double_buffered)))
        # Every step below is collective: a rank that fails still takes part in the exchanges of handles / verdicts, so
        # either all ranks end up with a working exchange or all of them raise (and callers fall back to NCCL together).
        ptr = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        err = None
        try:
            _lib.check(self.lib.tt_p2p_alloc(self.nbytes, C.byref(ptr)), "tt_p2p_alloc")
            _lib.check(self.lib.tt_p2p_export(ptr, handle), "tt_p2p_export")
        except RuntimeError as e:
            err = str(e)
        self.local_ptr = ptr.value
        handles = [None] * self.world
        dist.all_gather_object(handles, None if err else bytes(handle), group=group)
        self.desc = _lib.P2P()
        self.desc.world, self.desc.rank, self.desc.slot_bytes = self.world, self.rank, self.slot_bytes
        self.desc.double_buffered = int(self.double_buffered)
        # the kernel grid is a property of the exchange (same on every rank, every call): calls may then move any
        # byte count up to the slot and the arrival counters stay consistent
        self.desc.ctas = int(self.lib.tt_p2p_allgather_ctas(self.slot_bytes))
        self.desc.timeout_s = int(timeout_s or int(__import__("os").environ.get("TT_P2P_TIMEOUT_S", "0")))
        self._imported = []
        self._closed = False
        if err is None and all(h is not None for h in handles):
            try:
                for p in range(self.world):
                    if p == self.rank:
                        self.desc.base[p] = self.local_ptr
                        continue
                    h = (C.c_ubyte * 64).from_buffer_copy(handles[p])
                    q = C.c_void_p()
                    _lib.check(self.lib.tt_p2p_import(h, C.byref(q)), "tt_p2p_import")
                    self.desc.base[p] = q.value
                    self._imported.append(q.value)
            except RuntimeError as e:
                err = str(e)
        else:
            err = err or "a peer could not allocate / export its exchange buffer"
        verdicts = [None] * self.world
        dist.all_gather_object(verdicts, err, group=group)
        bad = [(r, v) for r, v in enumerate(verdicts) if v is not None]
        if bad:
            raise RuntimeError(f"peer-memory exchange unavailable (rank {bad[0][0]}: {bad[0][1]})")
        self._raw = torch.as_tensor(_DevMem(self.local_ptr, self.nbytes), device=self.device)   # uint8 view, not owning
        self.rounds = 0
        dist.barrier(group=group)                               # every peer has mapped every buffer before first use

    def gathered(self, dtype, per_rank_shape) -> torch.Tensor:
        """[world, *per_rank_shape] view of the slots (rank-major); slots are slot_bytes apart."""
        if self.double_buffered:
            raise RuntimeError("double-buffered exchanges are consumed through sum_slots()")
        n = 1
        for d in per_rank_shape:
            n *= int(d)
        esz = torch.empty((), dtype=dtype).element_size()
        if n * esz > self.slot_bytes:
            raise ValueError("per-rank shape exceeds the slot")
        body = self._raw[256:256 + self.world * self.slot_bytes].view(self.world, self.slot_bytes)
        return body[:, :n * esz].view(dtype).view(self.world, *per_rank_shape) if n * esz == self.slot_bytes else \
            torch.as_strided(body.view(dtype), (self.world, *per_rank_shape),
                             (self.slot_bytes // esz, *_contig_strides(per_rank_shape)))

    def current_half(self, round_no=None) -> torch.Tensor:
        """[world, slot_bytes] uint8 view of the slot set round `round_no` (default: the last allgather()) fills."""
        r = self.rounds if round_no is None else round_no
        off = 256 + (self.world * self.slot_bytes if (self.double_buffered and (r & 1)) else 0)
        return self._raw[off:off + self.world * self.slot_bytes].view(self.world, self.slot_bytes)

    def allgather(self, src: torch.Tensor, count: bool = True) -> None:
        """Launch the exchange on the current stream: afterwards slot r of the local buffer holds rank r's `src`.
        count=False while CAPTURING a CUDA graph (the kernel does not run yet): the caller bumps `rounds` per replay."""
        import ctypes as C
        from . import _lib
        if count:
            self.rounds += 1                                 # mirrors the device-side round counter (all calls go through here)
        nbytes = src.numel() * src.element_size()
        _lib.check(self.lib.tt_p2p_allgather(C.byref(self.desc), C.c_void_p(src.data_ptr()), nbytes,
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)), "tt_p2p_allgather")

    def check(self) -> None:
        """Synchronous: raise if an exchange kernel on this rank gave up waiting for a peer (its output was garbage)."""
        import ctypes as C
        from . import _lib
        who = C.c_int(-1)
        _lib.check(self.lib.tt_p2p_status(C.byref(self.desc), C.byref(who)), "tt_p2p_status")
        if who.value >= 0:
            raise RuntimeError(f"peer-memory exchange: rank {self.rank} timed out waiting for rank {who.value}")

    def close(self) -> None:
        """Unmap the peers' buffers and free the local one (every rank, after a barrier: nobody may still push)."""
        if self._closed:
            return
        self._closed = True
        try:
            torch.cuda.synchronize()
            for q in self._imported:
                self.lib.tt_p2p_unimport(q)
            if self.local_ptr:
                self.lib.tt_p2p_free(self.local_ptr)
        except Exception:
            pass
        self._imported, self.local_ptr, self._raw = [], None, None

    def sum_slots(self, out: torch.Tensor) -> None:
        """out[i] = sum over ranks (rank order) of the fp32 slots -- all-gather + this = a deterministic all-reduce."""
        import ctypes as C
        from . import _lib
        _lib.check(self.lib.tt_p2p_sum_slots(C.byref(self.desc), out.numel(), C.c_void_p(out.data_ptr()),
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)), "tt_p2p_sum_slots")


def _contig_strides(shape):
    st, acc = [], 1
    for d in reversed(shape):
        st.append(acc)
        acc *= int(d)
    return tuple(reversed(st))


_SEARCH_XCH = {}


def _search_exchange(nbytes: int, group, device, ctas: int = 0):
    """Cached double-buffered peer-memory exchange for the candidate records of the sharded search (None: use NCCL).
    ctas > 0 fixes the exchange's CTA count (one CTA per query for the fused merge + exchange + merge kernel)."""
    import os
    if device.type != "cuda" or os.environ.get("TT_P2P", "1") == "0":
        return None
    _, ws = world(group)
    if ws < 2 or ws > 8:
        return None
    key = (id(group), (nbytes + 255) // 256 * 256, device.index, int(ctas))
    if key not in _SEARCH_XCH:
        while len(_SEARCH_XCH) >= 4:                          # bounded: one exchange per (group, record size) in use
            old = _SEARCH_XCH.pop(next(iter(_SEARCH_XCH)))
            if old is not None:
                dist.barrier(group=group)
                old.close()
        try:
            x = P2PExchange(key[1], group, device, double_buffered=True)
            if ctas > 0:
                x.desc.ctas = int(ctas)                       # same on every rank: all of them pass the same nq
            x.staging = torch.zeros(1, key[1], dtype=torch.uint8, device=device)
            _SEARCH_XCH[key] = x
        except RuntimeError:
            _SEARCH_XCH[key] = None                           # no peer access on this machine: NCCL path
    return _SEARCH_XCH[key]


def shard_bounds(n: int, rank: int, ws: int) -> Tuple[int, int]:
    """Contiguous row range [lo, hi) of shard `rank` (rows split as evenly as possible)."""
    base, rem = divmod(n, ws)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedTopK:
    """Row-sharded exact top-k as ONE CUDA-graph replay per call: local scan + top-k, candidate exchange over NVLink peer
    memory, merge.  At 8 shards the scan of a 10 M x 256 index takes ~0.1-0.2 ms, so the launch / Python overhead of the
    eager chain (sharded_topk) would be most of a query; the graph removes it.  Collective: every rank calls it in step.
    The returned tensors are static buffers, valid until the next call.  Falls back to sharded_topk when the
    peer-memory exchange is unavailable."""

    def __init__(self, index_shard: torch.Tensor, k: int, id_offset: int, kernels, group=None, cosine: bool = True, nq: int = 1):
        self.index, self.k, self.id_offset, self.kernels, self.group, self.cosine, self.nq = \
            index_shard, k, id_offset, kernels, group, cosine, nq
        self.q = torch.zeros(nq, index_shard.shape[1], dtype=torch.float32, device=index_shard.device)
        self.graphs, self.outs = [None, None], [None, None]
        self.xch = None
        if index_shard.is_cuda and hasattr(kernels, "packed_topk_buffer") and index_shard.shape[0] >= k:
            _, self.nbytes, self.id_off = kernels.packed_topk_buffer(nq, k, index_shard.device)
            import os
            fused_ok = hasattr(kernels, "topk_scan_p2p") and nq <= 64 and os.environ.get("TT_SEARCH_FUSED", "1") != "0"
            self.xch = _search_exchange(self.nbytes, group, index_shard.device, ctas=nq if fused_ok else 0)
            # scan, then ONE kernel: block merge + candidate exchange over NVLink + final merge (tt_topk_scan_p2p)
            self.fused = bool(fused_ok and self.xch is not None and kernels.topk_scan_p2p_ok(self.xch, nq, k))
            if self.fused:
                self.ws = torch.empty(kernels.topk_scan_workspace_bytes(index_shard.shape[0], index_shard.shape[1], nq, k),
                                      dtype=torch.uint8, device=index_shard.device)

    def _body(self, round_no: int, count: bool):
        x = self.xch
        if getattr(self, "fused", False):
            if count:
                x.rounds += 1                                    # mirrors the device-side round counter
            return self.kernels.topk_scan_p2p(self.index, self.q, self.k, x, cosine=self.cosine, id_offset=self.id_offset,
                                              workspace=self.ws)
        s_view, i_view = self.kernels.packed_views(x.staging[0, :self.nbytes], self.nq, self.k, self.id_off)
        self.kernels.topk_scan(self.index, self.q, self.k, cosine=self.cosine, id_offset=self.id_offset, out=(s_view, i_view))
        x.allgather(x.staging[0], count=count)
        return self.kernels.topk_merge_packed(x.current_half(round_no), self.nq, self.k, self.id_off)

    def __call__(self, queries: torch.Tensor):
        if self.xch is None or queries.shape[0] != self.nq:
            return sharded_topk(self.index, queries, self.k, self.id_offset, self.kernels, self.group, self.cosine)
        self.q.copy_(queries)
        nxt = self.xch.rounds + 1
        par = nxt & 1
        if self.graphs[par] is None:
            out = self._body(nxt, count=True)                  # eager once for this parity (warms every kernel), ...
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):                          # ... then record the same chain for later replays
                self.outs[par] = self._body(nxt, count=False)
            self.graphs[par] = g
            return out
        self.xch.rounds = nxt
        self.graphs[par].replay()
        return self.outs[par]


def sharded_topk(index_shard: torch.Tensor, queries: torch.Tensor, k: int, id_offset: int, kernels, group=None,
                 cosine: bool = True):
    """Local scan + top-k on this rank's rows, all-gather the k candidates, merge.

    Every rank returns the same (scores [nq,k], global ids [nq,k]).  Shards smaller than k pad
    their lists with (-inf, -1), which the merge kernel ignores.
    """
    _, ws = world(group)
    nq = queries.shape[0]
    n_local = index_shard.shape[0]
    k_local = min(k, n_local)
    if ws == 1 and k_local == k:
        return kernels.topk_scan(index_shard, queries, k_local, cosine=cosine, id_offset=id_offset)
    packed = hasattr(kernels, "packed_topk_buffer")
    if packed:
        # (score, id) records of this rank go straight into one byte buffer -> ONE all-gather -> merge kernel
        mine, nbytes, id_off = kernels.packed_topk_buffer(nq, k, queries.device)
        xch = _search_exchange(nbytes, group, queries.device)
        if xch is not None:                                   # candidates travel over NVLink peer memory (one kernel)
            mine = xch.staging[:, :nbytes]
        s_view, i_view = kernels.packed_views(mine[0], nq, k, id_off)
        if k_local < k:
            s_view.fill_(float("-inf"))
            i_view.fill_(-1)
        if k_local == k:
            kernels.topk_scan(index_shard, queries, k, cosine=cosine, id_offset=id_offset, out=(s_view, i_view))
        elif k_local > 0:
            s, i = kernels.topk_scan(index_shard, queries, k_local, cosine=cosine, id_offset=id_offset)
            s_view[:, :k_local].copy_(s)
            i_view[:, :k_local].copy_(i)
        if xch is not None:
            xch.allgather(xch.staging[0])
            everyone = xch.current_half()                       # [R, slot_bytes]; records start each slot
        else:
            everyone = all_gather_rows(mine, group)             # [R, nbytes]
        return kernels.topk_merge_packed(everyone, nq, k, id_off)
    if k_local > 0:
        s, i = kernels.topk_scan(index_shard, queries, k_local, cosine=cosine, id_offset=id_offset)
    else:
        s = torch.empty(nq, 0, dtype=torch.float32, device=queries.device)
        i = torch.empty(nq, 0, dtype=torch.int64, device=queries.device)
    if k_local < k:
        pad_s = torch.full((nq, k - k_local), float("-inf"), dtype=torch.float32, device=s.device)
        pad_i = torch.full((nq, k - k_local), -1, dtype=torch.int64, device=s.device)
        s, i = torch.cat([s, pad_s], 1), torch.cat([i, pad_i], 1)
    all_s = all_gather_rows(s.unsqueeze(0), group)          # [R,nq,k]
    all_i = all_gather_rows(i.unsqueeze(0), group)
    return kernels.topk_merge(all_s, all_i)
