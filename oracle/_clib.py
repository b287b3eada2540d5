"""Compile and load oracle/vfi_oracle.c (plain C, gcc).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import hashlib
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "vfi_oracle.c"
OUT_DIR = HERE / "_build"
LIB = OUT_DIR / "libvfo.so"
STAMP = OUT_DIR / "libvfo.stamp"
# -ffp-contract=off: every fp32/fp64 operation rounds on its own, as written in the source.
# No -march flag: the .so travels to the GPU box, whose host CPU may differ.
CFLAGS = ["-O2", "-ffp-contract=off", "-fPIC", "-shared"]

_lib = None


def build(force: bool = False) -> Path:
    OUT_DIR.mkdir(exist_ok=True)
    digest = hashlib.sha256(SRC.read_bytes() + " ".join(CFLAGS).encode()).hexdigest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB
    proc = subprocess.run(["gcc", *CFLAGS, "-o", str(LIB), str(SRC), "-lm"], capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("gcc failed building the oracle:\n" + proc.stderr)
    STAMP.write_text(digest)
    return LIB


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(str(build()))
        P, i64, i32, f32 = C.c_void_p, C.c_int64, C.c_int, C.c_float
        l.vfo_bf16_round.argtypes = [P, i64]
        l.vfo_normalize_l2.argtypes = [P, i64, i32]
        l.vfo_canon_dot.argtypes = [P, P, i32]
        l.vfo_canon_dot.restype = f32
        l.vfo_rescore.argtypes = [P, P, i32, P, i64, P]
        l.vfo_topk.argtypes = [P, i64, i32, i64, P, P]
        l.vfo_topk_pairs.argtypes = [P, P, i64, i32, P, P]
        l.vfo_flat_search.argtypes = [P, i64, P, i64, i32, i32, i64, P, P]
        l.vfo_bm25_scores.argtypes = [P, P, P, i64, P, i64, i64, P]
        l.vfo_rrf.argtypes = [P, i32, i32, f32, i32, P, P]
        l.vfo_union.argtypes = [P, P, i32, i32, P, P, P]
        l.vfo_union.restype = i32
        for name in ("vfo_bf16_round", "vfo_normalize_l2", "vfo_rescore", "vfo_topk", "vfo_topk_pairs",
                     "vfo_flat_search", "vfo_bm25_scores", "vfo_rrf"):
            getattr(l, name).restype = None
        _lib = l
    return _lib


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)
