"""Brute-force two-tower retrieval behind the reference's ``BaseSearch`` interface.

Reference: ``BaseSearch`` ABC inference/search/base.py:8-54; ``TwoTowerSearch``
inference/search/two_tower.py:15-155.  Same constructor, attributes (``document_embeddings``,
``documents``), result schema (``{"document", "score"}`` sorted by descending score),
exceptions (``ValueError`` before indexing) and pickle format; the work is done by
  * index_documents: large-batch tokenise -> one H2D copy per batch -> document tower writing
    into ONE preallocated contiguous [N,H] fp32 (or bf16) matrix (reference: batches of 32 +
    torch.cat, two_tower.py:48-69);
  * search: query tower -> fused scan + exact top-k kernel (reference: F.cosine_similarity over
    a broadcast [1,N,H] + torch.topk, two_tower.py:98-105).
With a process group the index is row-sharded across ranks and candidates are merged with an
all-gather.
"""
from __future__ import annotations

import pickle
from abc import ABC, abstractmethod
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import ops, parallel


class BaseSearch(ABC):
    """inference/search/base.py:8-54"""

    @abstractmethod
    def index_documents(self, documents: List[str]) -> None:
        ...

    @abstractmethod
    def search(self, query: str, top_k: int = 5) -> List[Dict[str, Union[str, float]]]:
        ...

    def save_index(self, filepath: str) -> None:
        raise NotImplementedError("This search implementation does not support saving indices")

    def load_index(self, filepath: str) -> None:
        raise NotImplementedError("This search implementation does not support loading indices")


class TwoTowerSearch(BaseSearch):
    def __init__(self, model, tokenizer, device="cuda", index_dtype: str = "fp32", max_len: int = 64,
                 encode_batch_size: int = 8192, cosine: bool = True, process_group=None, kernels=None):
        if index_dtype not in ("fp32", "bf16"):
            raise ValueError("index_dtype must be 'fp32' or 'bf16'")
        self.model = model
        self.tokenizer = tokenizer
        self.device = device
        self.index_dtype = index_dtype
        self.max_len = max_len
        self.encode_batch_size = encode_batch_size
        self.cosine = cosine
        self.group = process_group
        self.kernels = kernels if kernels is not None else ops
        self.document_embeddings: Optional[torch.Tensor] = None      # this rank's rows
        self.documents: Optional[Sequence[str]] = None
        self.row_offset = 0                                          # global id of local row 0
        self.num_documents = 0
        self.batched_min_queries = 4                                 # search_batch: tensor-core scan from this batch size on
        self.model = self.model.to(self.device)

    # ------------------------------------------------------------------ indexing
    def _encode(self, texts: Sequence[str], tower) -> torch.Tensor:
        ids = self.tokenizer.encode_batch(texts, self.max_len, pin_memory=str(self.device).startswith("cuda"))
        ids = ids.to(self.device, non_blocking=True)
        with torch.no_grad():
            return tower(ids)

    def index_documents(self, documents: List[str]) -> None:
        self.documents = documents
        self.model.eval()
        n = len(documents)
        rank, ws = parallel.world(self.group) if self.group is not None or ws_initialized() else (0, 1)
        lo, hi = parallel.shard_bounds(n, rank, ws)
        self.row_offset, self.num_documents = lo, n
        hidden = self.model.document_tower.hidden_dim
        dtype = torch.float32 if self.index_dtype == "fp32" else torch.bfloat16
        mat = torch.empty(hi - lo, hidden, dtype=dtype, device=self.device)
        bs = self.encode_batch_size
        for i in range(lo, hi, bs):
            j = min(i + bs, hi)
            emb = self._encode(documents[i:j], self.model.document_tower)
            if dtype == torch.float32:
                mat[i - lo:j - lo].copy_(emb)
            else:
                self.kernels.cast_bf16(emb, mat[i - lo:j - lo])
        self.document_embeddings = mat

    def set_index(self, embeddings: torch.Tensor, documents: Sequence[str], row_offset: int = 0,
                  num_documents: Optional[int] = None) -> None:
        """Install a pre-computed [N,H] matrix (fp32 or bf16, this rank's rows) -- used for large
        synthetic indices where `documents` is a lazy sequence."""
        if embeddings.dim() != 2 or not embeddings.is_contiguous():
            raise ValueError("embeddings must be a contiguous [N,H] matrix")
        self.document_embeddings = embeddings
        self.documents = documents
        self.row_offset = row_offset
        self.num_documents = len(documents) if num_documents is None else num_documents
        self.index_dtype = "bf16" if embeddings.dtype == torch.bfloat16 else "fp32"

    # ------------------------------------------------------------------ search
    def _topk(self, q_emb: torch.Tensor, k: int):
        sharded = self.group is not None or ws_initialized()
        # batches on a bf16 index: ONE pass over the index per 128 queries on the tensor cores (tt_topk_scan_batched)
        # instead of one pass per 2 queries; the shards' candidate lists are merged like the single-query ones
        if (q_emb.shape[0] >= self.batched_min_queries and self.document_embeddings.is_cuda and
                k <= self.document_embeddings.shape[0] and hasattr(self.kernels, "topk_scan_batched_ok") and
                self.kernels.topk_scan_batched_ok(self.document_embeddings, k)):
            inv = None
            if self.cosine:
                key = self.document_embeddings.data_ptr()
                if getattr(self, "_inv_norms_key", None) != key:
                    self._inv_norms = self.kernels.index_row_inv_norms(self.document_embeddings)
                    self._inv_norms_key = key
                inv = self._inv_norms
            s, i = self.kernels.topk_scan_batched(self.document_embeddings, q_emb, k, id_offset=self.row_offset, row_inv_norms=inv)
            if not sharded:
                return s, i
            all_s = parallel.all_gather_rows(s.unsqueeze(0), self.group)
            all_i = parallel.all_gather_rows(i.unsqueeze(0), self.group)
            return self.kernels.topk_merge(all_s, all_i)
        if sharded:
            if self.document_embeddings.is_cuda and hasattr(parallel, "ShardedTopK"):
                key = (k, q_emb.shape[0], self.document_embeddings.data_ptr())
                cache = self.__dict__.setdefault("_sharded", {})
                if key not in cache:                            # one CUDA graph (scan, peer-memory exchange, merge) per shape
                    cache.clear()
                    cache[key] = parallel.ShardedTopK(self.document_embeddings, k, self.row_offset, self.kernels, self.group,
                                                      self.cosine, nq=q_emb.shape[0])
                s, i = cache[key](q_emb)
                return s.clone(), i.clone()                     # the graph's outputs are static buffers
            return parallel.sharded_topk(self.document_embeddings, q_emb, k, self.row_offset, self.kernels,
                                         self.group, self.cosine)
        return self.kernels.topk_scan(self.document_embeddings, q_emb, k, cosine=self.cosine,
                                      id_offset=self.row_offset)

    def search_batch(self, queries: Sequence[str], top_k: int = 5) -> List[List[Dict[str, Union[str, float]]]]:
        if self.document_embeddings is None:
            raise ValueError("No documents indexed. Call index_documents() first.")
        self.model.eval()
        k = min(top_k, self.num_documents)
        q_emb = self._encode(list(queries), self.model.query_tower)
        scores, ids = self._topk(q_emb, k)
        scores, ids = scores.cpu().tolist(), ids.cpu().tolist()
        return [[{"document": self.documents[i], "score": s} for s, i in zip(srow, irow)]
                for srow, irow in zip(scores, ids)]

    def search(self, query: str, top_k: int = 5) -> List[Dict[str, Union[str, float]]]:
        return self.search_batch([query], top_k)[0]

    # ------------------------------------------------------------------ persistence (two_tower.py:117-155)
    def save_index(self, filepath: str) -> None:
        if self.document_embeddings is None or self.documents is None:
            raise ValueError("No index to save. Call index_documents() first.")
        if self.row_offset != 0 or self.document_embeddings.shape[0] != self.num_documents:
            # the reference's pickle holds ONE matrix for ALL documents (two_tower.py:127-130); a rank of a row-sharded
            # index owns only its rows, and every rank would write the same path
            raise ValueError("save_index: this index is row-sharded across ranks; use save_index_raw(dir) per rank "
                             "(it records the row range) or build the index on one rank")
        emb = self.document_embeddings
        emb = emb.float().cpu().numpy() if torch.is_tensor(emb) else emb
        with open(filepath, "wb") as f:
            pickle.dump({"embeddings": emb, "documents": self.documents}, f)

    def load_index(self, filepath: str) -> None:
        with open(filepath, "rb") as f:
            data = pickle.load(f)
        emb = data["embeddings"]
        emb = torch.tensor(emb, device=self.device) if isinstance(emb, np.ndarray) else emb.to(self.device)
        emb = emb.contiguous()
        if self.index_dtype == "bf16" and emb.dtype == torch.float32:
            emb = self.kernels.cast_bf16(emb)
        self.document_embeddings = emb
        self.documents = data["documents"]
        self.row_offset, self.num_documents = 0, len(self.documents)


    # ------------------------------------------------------------------ raw, mmap-able format
    def save_index_raw(self, dirpath: str) -> None:
        """`dirpath/embeddings.bin` = the [N,H] matrix exactly as it sits in HBM (fp32 or bf16 bits, row-major),
        `dirpath/meta.json` = shape / dtype / row range, `dirpath/documents.json` = the documents.  A 10 M x 256 index is
        ONE sequential 10 GB (5 GB bf16) write instead of a pickled ndarray, and loads with np.memmap in row chunks
        (a shard reads only its own rows)."""
        import json, os
        if self.document_embeddings is None or self.documents is None:
            raise ValueError("No index to save. Call index_documents() first.")
        os.makedirs(dirpath, exist_ok=True)
        emb = self.document_embeddings
        n, h = emb.shape
        bits = emb.view(torch.int16) if emb.dtype == torch.bfloat16 else emb
        with open(os.path.join(dirpath, "embeddings.bin"), "wb") as f:
            for r0 in range(0, n, 1 << 20):                       # 1 M rows at a time: bounded host staging
                f.write(bits[r0:r0 + (1 << 20)].cpu().numpy().tobytes())
        with open(os.path.join(dirpath, "meta.json"), "w") as f:
            json.dump({"rows": n, "dim": h, "dtype": "bf16" if emb.dtype == torch.bfloat16 else "fp32",
                       "row_offset": self.row_offset, "num_documents": self.num_documents}, f)
        with open(os.path.join(dirpath, "documents.json"), "w") as f:
            json.dump(list(self.documents), f)

    def load_index_raw(self, dirpath: str, rows: Optional[range] = None) -> None:
        """Load `rows` (default: all) of an index written by save_index_raw straight into one device matrix."""
        import json, os
        with open(os.path.join(dirpath, "meta.json")) as f:
            meta = json.load(f)
        with open(os.path.join(dirpath, "documents.json")) as f:
            self.documents = json.load(f)
        n, h = meta["rows"], meta["dim"]
        lo, hi = (0, n) if rows is None else (rows.start, rows.stop)
        np_dt, t_dt = (np.int16, torch.bfloat16) if meta["dtype"] == "bf16" else (np.float32, torch.float32)
        mm = np.memmap(os.path.join(dirpath, "embeddings.bin"), dtype=np_dt, mode="r", shape=(n, h))
        out = torch.empty(hi - lo, h, dtype=t_dt, device=self.device)
        view = out.view(torch.int16) if t_dt == torch.bfloat16 else out
        for r0 in range(lo, hi, 1 << 20):
            r1 = min(hi, r0 + (1 << 20))
            view[r0 - lo:r1 - lo].copy_(torch.from_numpy(np.array(mm[r0:r1])), non_blocking=False)   # np.array: writable host copy of the chunk
        if self.index_dtype == "bf16" and t_dt == torch.float32:
            out = self.kernels.cast_bf16(out)
        elif self.index_dtype == "fp32" and t_dt == torch.bfloat16:
            out = out.float()
        self.document_embeddings = out
        self.row_offset, self.num_documents = meta.get("row_offset", 0) + lo, meta.get("num_documents", n)


def ws_initialized() -> bool:
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
