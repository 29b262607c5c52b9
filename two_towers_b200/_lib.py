"""ctypes binding of ``libtt_b200.so`` (the C ABI declared in ``include/tt_b200.h``).

The library is built in-tree by ``two_towers_b200/csrc/Makefile`` (``__graft_entry__.build()``).
There is NO fallback: if the shared object is missing, or a compute entry point is called
without an sm_100 device, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libtt_b200.so")

TT_PREC_FP32 = 0
TT_PREC_BF16 = 1
TT_TOPK_MAX = 1024
TT_ERR_UNSUPPORTED = -5
TT_ABI_VERSION = 6          # must equal TT_ABI_VERSION in include/tt_b200.h (a stale .so fails loudly at load)

_vp, _i, _i64, _f, _sz, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t, C.c_uint64

# name -> (restype, argtypes); mirrors include/tt_b200.h one to one
SIGNATURES = {
    "tt_abi_version": (_i, []),
    "tt_last_error": (C.c_char_p, []),
    "tt_require_sm100": (_i, [_i]),
    "tt_launch_count": (_i64, []),
    "tt_bad_token_id": (_i, [_vp, _i]),
    "tt_embed_gather": (_i, [_vp, _i, _vp, _i64, _i64, _i, _vp, _vp]),
    "tt_embed_pool_fwd": (_i, [_vp, _i, _vp, _i64, _i, _i64, _i, _vp, _vp, _vp, _vp, _vp]),
    "tt_embed_pool_bwd_workspace": (_sz, [_i64, _i, _i64, _i]),
    "tt_embed_pool_bwd": (_i, [_vp, _i, _vp, _vp, _i64, _i, _i64, _i, _vp, _vp, _sz, _vp]),
    "tt_mlp_workspace": (_sz, [_i64, _i, _i, _i]),
    "tt_mlp_fwd": (_i, [_vp] * 5 + [_i64, _i, _i] + [_vp] * 10 + [_i, _vp, _sz, _vp]),
    "tt_mlp_fwd_embed_ok": (_i, [_i, _i, _i64]),
    "tt_mlp_bwd": (_i, [_vp] * 6 + [_i64, _i, _i] + [_vp] * 9 + [_i, _i64, _vp, _vp, _vp, _vp, _vp] + [_i, _vp, _sz, _vp]),
    "tt_mlp_embed_workspace": (_sz, [_i64, _i, _i64]),
    "tt_proj_ln_workspace": (_sz, [_i64, _i, _i, _i]),
    "tt_proj_ln_fwd": (_i, [_vp] * 5 + [_i64, _i, _i, _i, _f, _i, _u64, _vp] + [_vp] * 4 + [_i, _vp, _sz, _vp]),
    "tt_proj_ln_bwd": (_i, [_vp] * 7 + [_i64, _i, _i, _i, _f, _i, _u64, _vp] + [_vp] * 5 + [_i, _vp, _sz, _vp]),
    "tt_inbatch_ce_workspace": (_sz, [_i64, _i64, _i, _i]),
    "tt_inbatch_ce_fwd": (_i, [_vp] * 4 + [_i64, _i64, _i, _f, _i64, _f] + [_vp] * 3 + [_i, _vp, _sz, _vp]),
    "tt_inbatch_ce_bwd": (_i, [_vp] * 5 + [_i64, _i64, _i, _f, _i64, _f] + [_vp] * 3 + [_i, _vp, _sz, _vp]),
    "tt_inbatch_ce_bwd_nparts": (_i, [_i64, _i64, _i, _i]),
    "tt_inbatch_ce_bwd_parts": (_i, [_vp] * 3 + [_i64, _i64, _i, _f, _i64, _f] + [_vp, _vp, _i64, _vp, _i64, _vp]),
    "tt_triplet_fwd": (_i, [_vp] * 3 + [_i64, _i, _f] + [_vp] * 4 + [_vp]),
    "tt_triplet_bwd": (_i, [_vp] * 4 + [_i64, _i, _f] + [_vp] * 4 + [_vp]),
    "tt_multineg_fwd": (_i, [_vp] * 3 + [_i64, _i, _i, _f] + [_vp] * 2 + [_vp]),
    "tt_multineg_bwd": (_i, [_vp] * 4 + [_i64, _i, _i, _f] + [_vp] * 4 + [_vp]),
    "tt_topk_scan_workspace": (_sz, [_i64, _i, _i, _i]),
    "tt_topk_scan": (_i, [_vp, _i, _vp, _i64, _i, _i, _i, _i, _i64, _vp, _vp, _vp, _sz, _vp]),
    "tt_topk_merge": (_i, [_vp, _vp, _i, _i, _i, _i64, _i64, _vp, _vp, _vp]),
    "tt_topk_scan_batched_ok": (_i, [_i, _i]),
    "tt_topk_scan_batched_workspace": (_sz, [_i64, _i, _i]),
    "tt_index_row_inv_norms": (_i, [_vp, _i64, _i, _vp, _vp]),
    "tt_topk_scan_batched": (_i, [_vp, _vp, _i64, _i, _i, _i, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "tt_cast_f32_to_bf16": (_i, [_vp, _vp, _i64, _vp]),
    "tt_selftest_tc_gemm": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _i, _vp, _vp]),
    "tt_adamw_step": (_i, [_vp] * 4 + [_i64] + [C.c_double] * 5 + [_vp, _vp, _vp]),
    "tt_adamw_step_publish": (_i, [_vp] * 4 + [_i64] + [C.c_double] * 5 + [_vp, _vp, _vp, _vp, _vp]),
    "tt_adamw_step_extra": (_i, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _i64] + [C.c_double] * 5 + [_vp, _vp, _vp, _vp, _vp]),
}

class MlpEmbed(C.Structure):
    """tt_mlp_embed_t (include/tt_b200.h)"""
    _fields_ = [("pool_bf16", _vp), ("V", _i64), ("table", _vp), ("table_bf16", _vp), ("d_table", _vp), ("accumulate", _i),
                ("workspace", _vp), ("workspace_bytes", _sz), ("ids", _vp), ("id_bytes", _i), ("L", _i), ("inv_len", _vp)]


class P2P(C.Structure):
    """tt_p2p_t (include/tt_b200.h)"""
    _fields_ = [("world", _i), ("rank", _i), ("slot_bytes", _sz), ("base", _vp * 8), ("double_buffered", _i),
                ("ctas", _i), ("timeout_s", _i)]


class CePass(C.Structure):
    """tt_ce_pass_t (include/tt_b200.h)"""
    _fields_ = [("x_bf16", _vp), ("x_rows", _i64), ("y_bf16", _vp), ("y_rows", _i64), ("y_buf_rows", _i64),
                ("y_blk", _i64), ("y_blk_stride", _i64), ("y_blk_off", _i64), ("lse", _vp), ("label_offset", _i64),
                ("out_parts", _vp), ("part_stride", _i64), ("dz_bf16", _vp), ("dz_colsum", _vp), ("inv_norm", _vp)]


SIGNATURES.update({
    "tt_inbatch_ce_fwd_ex_workspace": (_sz, [_i64, _i64]),
    "tt_inbatch_ce_fwd_ex": (_i, [_vp, _i64, _vp, _i64, _i64, _i64, _i64, _i64, _i, _f, _i64, _f, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "tt_inbatch_ce_sync_bytes": (_sz, [_i64]),
    "tt_p2p_buffer_bytes": (_sz, [_i, _sz, _i]),
    "tt_p2p_alloc": (_i, [_sz, _vp]),
    "tt_p2p_free": (_i, [_vp]),
    "tt_p2p_export": (_i, [_vp, _vp]),
    "tt_p2p_import": (_i, [_vp, _vp]),
    "tt_p2p_unimport": (_i, [_vp]),
    "tt_p2p_allgather_ctas": (_i, [_sz]),
    "tt_p2p_allgather": (_i, [_vp, _vp, _sz, _vp]),
    "tt_p2p_status": (_i, [_vp, _vp]),
    "tt_p2p_sum_slots": (_i, [_vp, _sz, _vp, _vp]),
    "tt_topk_scan_p2p_ok": (_i, [C.POINTER(P2P), _i, _i]),
    "tt_topk_scan_p2p": (_i, [_vp, _i, _vp, _i64, _i, _i, _i, _i, _i64, C.POINTER(P2P), _vp, _vp, _vp, _sz, _vp]),
    "tt_adamw_step_p2p": (_i, [_vp, _vp, C.POINTER(P2P), _vp, _vp, _i64] + [C.c_double] * 5 + [_vp, _vp, _vp, _vp, _vp]),
    "tt_inbatch_ce_bwd_fused_ok": (_i, [_i64, _i64, _i64, _i64, _i]),
    "tt_inbatch_ce_bwd_nparts_ex": (_i, [_i64, _i64, _i64, _i64, _i]),
    "tt_inbatch_ce_bwd_parts_ex": (_i, [C.POINTER(CePass), C.POINTER(CePass), _i, _f, _f, _vp, _i, _vp]),
    "tt_inbatch_ce_onepass_ok": (_i, [_i64, _i64, _i, _f]),
    "tt_inbatch_ce_onepass_sync_bytes": (_sz, [_i64]),
    "tt_inbatch_ce_fwd_dq": (_i, [C.POINTER(CePass), _i, _f, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tt_inbatch_ce_dd_nparts": (_i, [_i64, _i64, _i]),
    "tt_inbatch_ce_stash_ok": (_i, [_i64, _i64, _i]),
    "tt_inbatch_ce_stash_bytes": (_sz, [_i64, _i64, _i]),
    "tt_inbatch_ce_fwd_dq_stash": (_i, [C.POINTER(CePass), _i, _f, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tt_inbatch_ce_dd_stash": (_i, [C.POINTER(CePass), _i, _f, _f, _vp, _vp, _vp]),
    "tt_inbatch_ce_onepass": (_i, [C.POINTER(CePass), C.POINTER(CePass), _i, _f, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tt_inbatch_ce_fwd_dq_p2p": (_i, [C.POINTER(CePass), _i, _f, _f, _f, _vp, _vp, _vp, _vp, _vp, C.POINTER(P2P), _vp, _vp]),
    "tt_inbatch_ce_dd": (_i, [C.POINTER(CePass), _i, _f, _f, _vp, _vp]),
})

_lib = None


def load():
    """Load libtt_b200.so (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"two_towers_b200: {LIB_PATH} is missing -- build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C two_towers_b200/csrc`). "
            "There is no CPU / PyTorch fallback for the hot path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError == ABI mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.tt_abi_version() != TT_ABI_VERSION:
        raise RuntimeError(f"two_towers_b200: ABI version {lib.tt_abi_version()} != {TT_ABI_VERSION} (rebuild: make -C two_towers_b200/csrc)")
    _lib = lib
    return lib


def last_error() -> str:
    return load().tt_last_error().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")


def launch_count() -> int:
    return int(load().tt_launch_count())


def raise_on_bad_ids() -> None:
    """IndexError if a completed gather kernel has seen a token id outside [0, V) -- where the reference's nn.Embedding
    raises (twotower/embeddings.py:33-40).  Call after the synchronisation that made the results visible."""
    bad = C.c_int64(0)
    if load().tt_bad_token_id(C.byref(bad), 1):
        raise IndexError(f"index out of range in self (token id {bad.value} is outside the embedding table)")
