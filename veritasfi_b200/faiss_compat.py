"""Drop-in for the slice of the `faiss` module the reference uses.

Reference call sites (paths relative to /root/reference/):
    faiss.IndexFlatIP(dimension)            src/utils/faissRetriever.py:18
    faiss.normalize_L2(x)                   src/utils/faissRetriever.py:22,35
    index.add(x)                            src/utils/faissRetriever.py:24
    index.search(query_vector, k) -> (D, I) src/utils/faissRetriever.py:37

Same names, argument meaning and error behaviour as faiss-cpu [upstream]: float32 C-contiguous
2-D inputs, `search` returns (distances float32 [nq,k], labels int64 [nq,k]) as host numpy arrays,
rows short of k padded with label -1 / distance -FLT_MAX, dimension mismatch raises.
Everything runs on the GPU through the C ABI (include/vfi.h); there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N

METRIC_INNER_PRODUCT = 0
_FLT_MAX = np.finfo(np.float32).max


def _as_f32_matrix(x, what: str) -> np.ndarray:
    if not isinstance(x, np.ndarray):
        raise TypeError(f"{what}: expected a numpy array, got {type(x).__name__}")
    if x.dtype != np.float32:
        raise TypeError(f"{what}: array must be float32 (got {x.dtype})")
    if x.ndim != 2:
        raise ValueError(f"{what}: array must be 2-dimensional")
    if not x.flags.c_contiguous:
        raise ValueError(f"{what}: array must be C-contiguous")
    return x


def normalize_L2(x: np.ndarray, device: int | None = None) -> None:
    """In-place row-wise L2 normalisation (zero rows untouched) — faiss.normalize_L2."""
    import os
    device = int(os.environ.get("VFI_DEVICE", "0")) if device is None else int(device)
    x = _as_f32_matrix(x, "normalize_L2")
    if not x.flags.writeable:
        raise ValueError("normalize_L2: array is read-only")
    n, d = x.shape
    if n == 0:
        return
    N.check(N.load().vfi_normalize_l2(x.ctypes.data_as(C.c_void_p), n, d, N.MEM_HOST, device, None))


class IndexFlatIP:
    """Exact inner-product index over fp32 rows held in HBM (faiss.IndexFlatIP)."""

    metric_type = METRIC_INNER_PRODUCT
    is_trained = True

    def __init__(self, d: int, device: int | None = None, store: str | None = None):
        # the reference constructs `faiss.IndexFlatIP(dimension)` (faissRetriever.py:18): where the index lives and how rows
        # are stored come from the environment when the caller does not say (VFI_DEVICE, default 0; VFI_STORE, default f32 =
        # faiss semantics on the fp32 values, bf16 = the rows are defined as their bf16 roundings)
        import os
        self._h = C.c_void_p()
        self.d = int(d)
        self.device = int(os.environ.get("VFI_DEVICE", "0")) if device is None else int(device)
        store = os.environ.get("VFI_STORE", "f32") if store is None else store
        store_code = {"f32": N.STORE_F32, "bf16": N.STORE_BF16}[store]
        self._store_code = store_code
        N.check(N.load().vfi_index_create(self.d, store_code, self.device, C.byref(self._h)))

    # -- faiss surface ---------------------------------------------------------------------
    @property
    def ntotal(self) -> int:
        return int(N.load().vfi_index_ntotal(self._h))

    def add(self, x: np.ndarray) -> None:
        x = _as_f32_matrix(x, "add")
        if x.shape[1] != self.d:
            raise AssertionError(f"add: vectors have dimension {x.shape[1]}, index has {self.d}")
        N.check(N.load().vfi_index_add(self._h, x.ctypes.data_as(C.c_void_p), x.shape[0], N.MEM_HOST, None))

    def search(self, x: np.ndarray, k: int):
        x = _as_f32_matrix(x, "search")
        if x.shape[1] != self.d:
            raise AssertionError(f"search: queries have dimension {x.shape[1]}, index has {self.d}")
        k = int(k)
        if k <= 0:
            raise AssertionError("search: k must be positive")
        nq = x.shape[0]
        D = np.full((nq, k), -_FLT_MAX, dtype=np.float32)
        I = np.full((nq, k), -1, dtype=np.int64)
        if nq:
            N.check(N.load().vfi_index_search(self._h, x.ctypes.data_as(C.c_void_p), nq, k,
                                              D.ctypes.data_as(C.c_void_p), I.ctypes.data_as(C.c_void_p),
                                              N.MEM_HOST, None))
        return D, I

    def reconstruct(self, i: int) -> np.ndarray:
        out = np.empty(self.d, dtype=np.float32)
        N.check(N.load().vfi_index_reconstruct(self._h, int(i), out.ctypes.data_as(C.c_void_p), N.MEM_HOST))
        return out

    def reset(self) -> None:
        self.close()
        N.check(N.load().vfi_index_create(self.d, self._store_code, self.device, C.byref(self._h)))

    # -- extensions ------------------------------------------------------------------------
    def set_option(self, opt: int, value: int) -> None:
        N.check(N.load().vfi_index_set_option(self._h, opt, int(value)))

    def stats(self, reset: bool = False) -> N.SearchStats:
        st = N.SearchStats()
        N.check(N.load().vfi_index_get_stats(self._h, C.byref(st), int(reset)))
        return st

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            N.load().vfi_index_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
