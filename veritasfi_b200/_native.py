"""ctypes binding of include/vfi.h.  The product path: fails loudly when libvfi.so is missing or
when there is no CUDA device — there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

from .build import LIB_PATH

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_UNSUPPORTED, ERR_NO_DEVICE, ERR_INTERNAL = range(7)
MEM_HOST, MEM_DEVICE = 0, 1
STORE_BF16, STORE_F32 = 0, 1
DTYPE_F32, DTYPE_BF16 = 0, 1
OPT_OVERFETCH, OPT_FORCE_PATH, OPT_PROFILE, OPT_TAU_HINT, OPT_NUM_CTAS, OPT_CTA_PAIR, OPT_TAU_M, OPT_SMALL_BATCH = 1, 2, 3, 4, 5, 7, 9, 10
OPT_TAIL_PIECE = 11
PATH_AUTO, PATH_EXACT, PATH_FUSED, PATH_GEMV = 0, 1, 2, 3
PATH_EXHAUSTIVE = PATH_EXACT   # the old name: every row scored canonically
MAX_K = 2048
IPC_HANDLE_BYTES = 64


class VfiError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"vfi error {code}: {msg}")
        self.code = code


class NoDeviceError(VfiError):
    pass


class SearchStats(C.Structure):
    _fields_ = [
        ("searches", C.c_int64), ("queries", C.c_int64), ("retried_queries", C.c_int64),
        ("fused_launches", C.c_int64), ("fused_ms_total", C.c_double), ("fused_ms_samples", C.c_int64),
        ("last_path", C.c_int), ("last_overfetch", C.c_int), ("last_eps", C.c_float), ("max_abs_err", C.c_float),
        ("hint_retries", C.c_int64), ("tail_ms_total", C.c_double), ("tail_ms_samples", C.c_int64),
    ]


class Bm25Stats(C.Structure):
    _fields_ = [("launches", C.c_int64), ("score_ms_total", C.c_double), ("score_ms_samples", C.c_int64),
                ("postings_bytes", C.c_int64)]


_P = C.c_void_p
_SIGNATURES = {
    "vfi_abi_version": (C.c_int, []),
    "vfi_last_error": (C.c_char_p, []),
    "vfi_device_count": (C.c_int, []),
    "vfi_launch_count": (C.c_int64, []),
    "vfi_index_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "vfi_index_destroy": (C.c_int, [_P]),
    "vfi_index_reserve": (C.c_int, [_P, C.c_int64]),
    "vfi_index_add": (C.c_int, [_P, _P, C.c_int64, C.c_int, _P]),
    "vfi_index_add_bf16": (C.c_int, [_P, _P, C.c_int64, C.c_int, _P]),
    "vfi_index_ntotal": (C.c_int64, [_P]),
    "vfi_index_dim": (C.c_int, [_P]),
    "vfi_index_set_id_offset": (C.c_int, [_P, C.c_int64]),
    "vfi_index_reconstruct": (C.c_int, [_P, C.c_int64, _P, C.c_int]),
    "vfi_index_read_rows": (C.c_int, [_P, C.c_int64, C.c_int64, _P, C.c_int, _P]),
    "vfi_index_pairwise": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, _P]),
    "vfi_index_search": (C.c_int, [_P, _P, C.c_int64, C.c_int, _P, _P, C.c_int, _P]),
    "vfi_index_search_ex": (C.c_int, [_P, _P, C.c_int, C.c_int64, C.c_int, _P, _P, C.c_int, _P]),
    "vfi_index_search_begin_ex": (C.c_int, [_P, _P, C.c_int, C.c_int64, C.c_int, _P, _P, _P, C.POINTER(C.c_int)]),
    "vfi_index_set_option": (C.c_int, [_P, C.c_int, C.c_int64]),
    "vfi_index_get_stats": (C.c_int, [_P, C.POINTER(SearchStats), C.c_int]),
    "vfi_index_debug_scores": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int, _P]),
    "vfi_normalize_l2": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int, C.c_int, _P]),
    "vfi_cosine_topk": (C.c_int, [_P, C.c_int64, _P, C.c_int64, C.c_int, C.c_int, _P, _P, C.c_int, C.c_int, _P]),
    "vfi_merge_topk": (C.c_int, [_P, _P, C.c_int, C.c_int64, C.c_int, C.c_int, _P, _P, C.c_int, C.c_int, _P]),
    "vfi_exchange_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.POINTER(_P)]),
    "vfi_exchange_handle": (C.c_int, [_P, _P]),
    "vfi_exchange_connect": (C.c_int, [_P, _P]),
    "vfi_exchange_set_timeout_ms": (C.c_int, [_P, C.c_int64]),
    "vfi_exchange_merge": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int, C.c_int, _P, _P, _P]),
    "vfi_exchange_destroy": (C.c_int, [_P]),
    "vfi_bm25_create": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.POINTER(_P)]),
    "vfi_bm25_destroy": (C.c_int, [_P]),
    "vfi_bm25_ndocs": (C.c_int64, [_P]),
    "vfi_bm25_search": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int, _P, _P, C.c_int, _P]),
    "vfi_bm25_score_all": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int, _P]),
    "vfi_bm25_rank_all": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P]),
    "vfi_bm25_set_profile": (C.c_int, [_P, C.c_int]),
    "vfi_bm25_get_stats": (C.c_int, [_P, C.POINTER(Bm25Stats), C.c_int]),
    "vfi_fuse_rrf": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_int, _P, _P, C.c_int, C.c_int, _P]),
    "vfi_index_search_begin": (C.c_int, [_P, _P, C.c_int64, C.c_int, _P, _P, _P, C.POINTER(C.c_int)]),
    "vfi_index_search_finish": (C.c_int, [_P, C.c_int]),
    "vfi_stem_english": (C.c_int, [C.c_char_p, _P, C.c_int64, _P, C.c_int64, _P]),
    "vfi_tokenize_ascii": (C.c_int, [C.c_char_p, C.c_int64, _P, _P, C.c_int64, _P]),
    "vfi_fuse_union": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "vfi_fuse_hybrid": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int64, _P, C.c_int64, C.c_int, C.c_int,
                                  C.c_float, C.c_int, _P, _P, C.c_int, _P]),
    "vfi_index_ticket_flag": (C.c_int, [_P, C.c_int, C.POINTER(_P)]),
    "vfi_exchange_merge_flagged": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P]),
    "vfi_index_search_begin_push": (C.c_int, [_P, _P, C.c_int, C.c_int64, C.c_int, _P, _P, _P, _P, C.POINTER(C.c_int)]),
    "vfi_exchange_merge_pushed": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int, _P, _P, _P, C.POINTER(C.c_int), _P]),
    "vfi_exchange_any_fail": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int)]),
    "vfi_bm25_create_from": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.POINTER(_P)]),
    "vfi_bm25_rank_range": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int64, _P, _P, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def lib_path() -> Path:
    return Path(os.environ.get("VFI_LIB", str(LIB_PATH)))


def load():
    """dlopen libvfi.so and declare every prototype.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not path.exists():
        raise RuntimeError(
            f"{path} is missing: build it with `python -m veritasfi_b200.build` (nvcc, sm_100a). "
            "veritasfi_b200 has no CPU fallback.")
    lib = C.CDLL(str(path))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.vfi_abi_version() != 1:
        raise RuntimeError("libvfi.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


def check(code: int) -> None:
    if code == OK:
        return
    msg = (load().vfi_last_error() or b"").decode("utf-8", "replace")
    if code == ERR_NO_DEVICE:
        raise NoDeviceError(code, msg)
    raise VfiError(code, msg)


def device_count() -> int:
    return int(load().vfi_device_count())


def launch_count() -> int:
    return int(load().vfi_launch_count())
