"""Torch-facing dense index: the batch entry `search_batch(Q, k) -> (ids, scores)` of SURVEY.md §8b.

Device tensors go straight through the C ABI by pointer (no host round trip); torch is only the
owner of device memory and streams here.  Reference seam: FaissRetriever.invoke
(/root/reference/src/utils/faissRetriever.py:28-38) returns (indices, distances) — ids first."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as N

_FLT_MAX = float(np.finfo(np.float32).max)


def _stream_ptr(device: torch.device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class SearchTicket:
    """One batch in flight: keeps the query and result tensors alive until `DenseIndex.search_finish`."""
    __slots__ = ("ticket", "q", "ids", "scores")

    def __init__(self, ticket: int, q: torch.Tensor, ids: torch.Tensor, scores: torch.Tensor):
        self.ticket, self.q, self.ids, self.scores = ticket, q, ids, scores


class DenseIndex:
    """Flat inner-product index over a corpus shard resident in HBM.

    store="bf16": the corpus (and the queries) are defined as their bf16 roundings — BASELINE
    configs 2/3/5.  store="f32": faiss.IndexFlatIP semantics on fp32 values."""

    def __init__(self, d: int, store: str = "bf16", device: int | torch.device = 0, id_offset: int = 0):
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.d = int(d)
        self.store = store
        self._h = C.c_void_p()
        code = {"bf16": N.STORE_BF16, "f32": N.STORE_F32}[store]
        N.check(N.load().vfi_index_create(self.d, code, self.device.index or 0, C.byref(self._h)))
        if id_offset:
            self.set_id_offset(id_offset)

    @property
    def ntotal(self) -> int:
        return int(N.load().vfi_index_ntotal(self._h))

    def set_id_offset(self, off: int) -> None:
        N.check(N.load().vfi_index_set_id_offset(self._h, int(off)))

    def reserve(self, n: int) -> None:
        N.check(N.load().vfi_index_reserve(self._h, int(n)))

    def set_option(self, opt: int, value: int) -> None:
        N.check(N.load().vfi_index_set_option(self._h, opt, int(value)))

    def stats(self, reset: bool = False) -> N.SearchStats:
        st = N.SearchStats()
        N.check(N.load().vfi_index_get_stats(self._h, C.byref(st), int(reset)))
        return st

    def add(self, x) -> None:
        """Append rows.  x: numpy float32 [n,d] (host), or torch float32/bfloat16 [n,d] (cuda or cpu)."""
        lib = N.load()
        if isinstance(x, np.ndarray):
            if x.dtype != np.float32 or x.ndim != 2 or x.shape[1] != self.d or not x.flags.c_contiguous:
                raise ValueError("add: need a C-contiguous float32 [n, d] array")
            N.check(lib.vfi_index_add(self._h, x.ctypes.data_as(C.c_void_p), x.shape[0], N.MEM_HOST, None))
            return
        if not isinstance(x, torch.Tensor) or x.dim() != 2 or x.shape[1] != self.d:
            raise ValueError("add: need a [n, d] tensor")
        x = x.contiguous()
        mem = N.MEM_DEVICE if x.is_cuda else N.MEM_HOST
        if x.is_cuda and x.device != self.device:
            raise ValueError("add: tensor is on another device")
        stream = _stream_ptr(self.device) if x.is_cuda else None
        if x.dtype == torch.float32:
            N.check(lib.vfi_index_add(self._h, C.c_void_p(x.data_ptr()), x.shape[0], mem, stream))
        elif x.dtype == torch.bfloat16:
            N.check(lib.vfi_index_add_bf16(self._h, C.c_void_p(x.data_ptr()), x.shape[0], mem, stream))
        else:
            raise ValueError("add: dtype must be float32 or bfloat16")

    @staticmethod
    def _q_dtype(q: torch.Tensor) -> int:
        return N.DTYPE_BF16 if q.dtype == torch.bfloat16 else N.DTYPE_F32

    def search_batch(self, q: torch.Tensor, k: int):
        """q: float32 or bfloat16 [B,d] on this index's GPU.  Returns (ids int64 [B,k], scores float32 [B,k]) on the GPU."""
        if not (isinstance(q, torch.Tensor) and q.is_cuda and q.dtype in (torch.float32, torch.bfloat16) and q.dim() == 2
                and q.shape[1] == self.d):
            raise ValueError("search_batch: need a float32 or bfloat16 [B, d] cuda tensor")
        q = q.contiguous()
        B = q.shape[0]
        ids = torch.empty((B, k), dtype=torch.int64, device=self.device)
        scores = torch.empty((B, k), dtype=torch.float32, device=self.device)
        if B:
            N.check(N.load().vfi_index_search_ex(self._h, C.c_void_p(q.data_ptr()), self._q_dtype(q), B, int(k),
                                                 C.c_void_p(scores.data_ptr()), C.c_void_p(ids.data_ptr()),
                                                 N.MEM_DEVICE, _stream_ptr(self.device)))
        return ids, scores

    def search_begin(self, q: torch.Tensor, k: int, out=None, push=None) -> "SearchTicket":
        """Enqueue one batch (B <= 1024) and return at once; `search_finish(ticket)` waits, certifies and hands out
        (ids, scores).  Beginning batch i+1 before finishing batch i keeps the GPU busy while the host looks at the
        certificate flag of batch i (vfi_index_search_begin / vfi_index_search_finish).  out: optional contiguous
        (ids int64 [B,k], scores float32 [B,k]) tensors to write into.  push: a sharded.PeerExchange — the batch's rescoring
        kernel then also sends every finished row to the peers' windows (vfi_index_search_begin_push); the caller follows up
        with push.merge_pushed(...) on the same stream."""
        if not (isinstance(q, torch.Tensor) and q.is_cuda and q.dtype in (torch.float32, torch.bfloat16) and q.dim() == 2
                and q.shape[1] == self.d and q.shape[0] > 0):
            raise ValueError("search_begin: need a non-empty float32 or bfloat16 [B, d] cuda tensor")
        q = q.contiguous()
        B = q.shape[0]
        if out is None:
            ids = torch.empty((B, k), dtype=torch.int64, device=self.device)
            scores = torch.empty((B, k), dtype=torch.float32, device=self.device)
        else:
            ids, scores = out
            if not (ids.is_contiguous() and scores.is_contiguous() and ids.shape == (B, k) and scores.shape == (B, k)
                    and ids.dtype == torch.int64 and scores.dtype == torch.float32):
                raise ValueError("search_begin: out must be contiguous (int64 [B,k], float32 [B,k])")
        t = C.c_int(-1)
        if push is not None:
            N.check(N.load().vfi_index_search_begin_push(self._h, C.c_void_p(q.data_ptr()), self._q_dtype(q), B, int(k),
                                                         C.c_void_p(scores.data_ptr()), C.c_void_p(ids.data_ptr()), push._h,
                                                         _stream_ptr(self.device), C.byref(t)))
        else:
            N.check(N.load().vfi_index_search_begin_ex(self._h, C.c_void_p(q.data_ptr()), self._q_dtype(q), B, int(k),
                                                       C.c_void_p(scores.data_ptr()), C.c_void_p(ids.data_ptr()),
                                                       _stream_ptr(self.device), C.byref(t)))
        return SearchTicket(t.value, q, ids, scores)

    def search_finish(self, ticket: "SearchTicket"):
        N.check(N.load().vfi_index_search_finish(self._h, ticket.ticket))
        return ticket.ids, ticket.scores

    def ticket_flag_ptr(self, ticket: "SearchTicket") -> int:
        """Device address of the batch's certificate counter (vfi_index_ticket_flag); valid until search_finish."""
        p = C.c_void_p()
        N.check(N.load().vfi_index_ticket_flag(self._h, ticket.ticket, C.byref(p)))
        return p.value

    def search_host(self, q: np.ndarray, k: int):
        """Host in, host out (pinned or pageable numpy): the reference-facing call, copies included."""
        if q.dtype != np.float32 or q.ndim != 2 or q.shape[1] != self.d or not q.flags.c_contiguous:
            raise ValueError("search_host: need a C-contiguous float32 [B, d] array")
        B = q.shape[0]
        scores = np.full((B, k), -_FLT_MAX, dtype=np.float32)
        ids = np.full((B, k), -1, dtype=np.int64)
        if B:
            N.check(N.load().vfi_index_search(self._h, q.ctypes.data_as(C.c_void_p), B, int(k),
                                              scores.ctypes.data_as(C.c_void_p), ids.ctypes.data_as(C.c_void_p),
                                              N.MEM_HOST, None))
        return ids, scores

    def search_host_into(self, q_ptr: int, B: int, k: int, scores_ptr: int, ids_ptr: int, stream: int | None = None) -> None:
        """Same as search_host on caller-owned (e.g. pinned) buffers given by address.  stream: a cudaStream_t value; None =
        the call's own stream from the workspace pool (concurrent callers overlap)."""
        N.check(N.load().vfi_index_search(self._h, C.c_void_p(q_ptr), B, int(k), C.c_void_p(scores_ptr),
                                          C.c_void_p(ids_ptr), N.MEM_HOST, C.c_void_p(stream) if stream else None))

    # -- persist / reload a shard -------------------------------------------------
    def read_rows(self, first: int = 0, n: int | None = None) -> np.ndarray:
        """Stored rows [first, first+n) as float32 (bf16 rows are exactly representable)."""
        n = self.ntotal - first if n is None else n
        out = np.empty((n, self.d), dtype=np.float32)
        if n:
            N.check(N.load().vfi_index_read_rows(self._h, int(first), int(n), out.ctypes.data_as(C.c_void_p), N.MEM_HOST, None))
        return out

    def save(self, path: str, chunk: int = 1 << 18) -> None:
        """<path>.json (d, store, ntotal, id_offset) + <path>.rows.npy (float32 rows; reloading is bit-exact)."""
        import json
        n = self.ntotal
        mm = np.lib.format.open_memmap(path + ".rows.npy", mode="w+", dtype=np.float32, shape=(n, self.d))
        for r0 in range(0, n, chunk):
            r1 = min(n, r0 + chunk)
            mm[r0:r1] = self.read_rows(r0, r1 - r0)
        mm.flush()
        del mm
        with open(path + ".json", "w") as f:
            json.dump({"d": self.d, "store": self.store, "ntotal": n, "format": "vfi-dense-shard-1"}, f)

    @classmethod
    def load(cls, path: str, device: int | torch.device = 0, id_offset: int = 0, chunk: int = 1 << 18) -> "DenseIndex":
        import json
        with open(path + ".json") as f:
            meta = json.load(f)
        rows = np.load(path + ".rows.npy", mmap_mode="r")
        self = cls(meta["d"], store=meta["store"], device=device, id_offset=id_offset)
        self.reserve(meta["ntotal"])
        for r0 in range(0, meta["ntotal"], chunk):
            self.add(np.ascontiguousarray(rows[r0:r0 + chunk]))
        return self

    # -- similarity between stored rows, by id -------------------------------------
    def pairwise(self, ids) -> torch.Tensor:
        """out[i, j] = canonical <row ids[i], row ids[j]> on the GPU; with L2-normalised rows this is the cosine
        matrix of ensembleRetriever.py:265-281 without re-embedding any chunk text."""
        ids_t = torch.as_tensor(ids, dtype=torch.int64, device=self.device).contiguous()
        n = ids_t.numel()
        out = torch.empty((n, n), dtype=torch.float32, device=self.device)
        if n:
            N.check(N.load().vfi_index_pairwise(self._h, C.c_void_p(ids_t.data_ptr()), n, C.c_void_p(out.data_ptr()),
                                                N.MEM_DEVICE, _stream_ptr(self.device)))
        return out

    def debug_scores(self, q: torch.Tensor) -> torch.Tensor:
        """Raw tensor-core scores [B, ntotal] (test hook for the tcgen05 path)."""
        q = q.contiguous()
        out = torch.empty((q.shape[0], self.ntotal), dtype=torch.float32, device=self.device)
        N.check(N.load().vfi_index_debug_scores(self._h, C.c_void_p(q.data_ptr()), q.shape[0],
                                                C.c_void_p(out.data_ptr()), N.MEM_DEVICE, _stream_ptr(self.device)))
        return out

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            N.load().vfi_index_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def merge_topk(scores: torch.Tensor, ids: torch.Tensor, k_out: int):
    """scores float32 [G,B,k], ids int64 [G,B,k] on one GPU -> (ids [B,k_out], scores [B,k_out])."""
    G, B, k_in = scores.shape
    dev = scores.device
    out_s = torch.empty((B, k_out), dtype=torch.float32, device=dev)
    out_i = torch.empty((B, k_out), dtype=torch.int64, device=dev)
    N.check(N.load().vfi_merge_topk(C.c_void_p(scores.contiguous().data_ptr()), C.c_void_p(ids.contiguous().data_ptr()),
                                    G, B, k_in, int(k_out), C.c_void_p(out_s.data_ptr()), C.c_void_p(out_i.data_ptr()),
                                    N.MEM_DEVICE, dev.index or 0, _stream_ptr(dev)))
    return out_i, out_s
