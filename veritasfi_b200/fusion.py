"""Rank fusion on the GPU (K5).  Host numpy in/out or torch cuda tensors in/out.

union: the reference's shared-`seen_ids` ordered de-dup union across the FAISS, Title-Summary and
BM25 sections (/root/reference/src/utils/ensembleRetriever.py:58,72-74,148-150,194-196).
rrf:   reciprocal-rank fusion with k_rrf = 60 (north_star)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N

_FLT_MAX = np.finfo(np.float32).max


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def rrf(ids, k: int, k_rrf: float = 60.0, device: int = 0):
    """ids int64 [B,P,L] (-1 = padding).  Returns (ids [B,k], fused scores float32 [B,k])."""
    lib = N.load()
    if _is_torch(ids):
        import torch
        ids = ids.contiguous()
        B, P, L = ids.shape
        oi = torch.empty((B, k), dtype=torch.int64, device=ids.device)
        os_ = torch.empty((B, k), dtype=torch.float32, device=ids.device)
        st = C.c_void_p(torch.cuda.current_stream(ids.device).cuda_stream)
        N.check(lib.vfi_fuse_rrf(C.c_void_p(ids.data_ptr()), B, P, L, float(k_rrf), int(k), C.c_void_p(os_.data_ptr()),
                                 C.c_void_p(oi.data_ptr()), N.MEM_DEVICE, ids.device.index or 0, st))
        return oi, os_
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    B, P, L = ids.shape
    oi = np.full((B, k), -1, np.int64)
    os_ = np.full((B, k), -_FLT_MAX, np.float32)
    if B:
        N.check(lib.vfi_fuse_rrf(ids.ctypes.data_as(C.c_void_p), B, P, L, float(k_rrf), int(k),
                                 os_.ctypes.data_as(C.c_void_p), oi.ctypes.data_as(C.c_void_p), N.MEM_HOST, device, None))
    return oi, os_


def union(ids, scores, device: int = 0):
    """ids int64 [B,P,L], scores float32 [B,P,L].  Returns (ids [B,P*L], scores, path int32, count int32 [B]);
    entries past count are -1 / -FLT_MAX / -1."""
    lib = N.load()
    if _is_torch(ids):
        import torch
        ids, scores = ids.contiguous(), scores.contiguous()
        B, P, L = ids.shape
        dev = ids.device
        oi = torch.empty((B, P * L), dtype=torch.int64, device=dev)
        os_ = torch.empty((B, P * L), dtype=torch.float32, device=dev)
        op = torch.empty((B, P * L), dtype=torch.int32, device=dev)
        oc = torch.empty((B,), dtype=torch.int32, device=dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        N.check(lib.vfi_fuse_union(C.c_void_p(ids.data_ptr()), C.c_void_p(scores.data_ptr()), B, P, L,
                                   C.c_void_p(oi.data_ptr()), C.c_void_p(os_.data_ptr()), C.c_void_p(op.data_ptr()),
                                   C.c_void_p(oc.data_ptr()), N.MEM_DEVICE, dev.index or 0, st))
        return oi, os_, op, oc
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    B, P, L = ids.shape
    oi = np.full((B, P * L), -1, np.int64)
    os_ = np.full((B, P * L), -_FLT_MAX, np.float32)
    op = np.full((B, P * L), -1, np.int32)
    oc = np.zeros(B, np.int32)
    if B:
        N.check(lib.vfi_fuse_union(ids.ctypes.data_as(C.c_void_p), scores.ctypes.data_as(C.c_void_p), B, P, L,
                                   oi.ctypes.data_as(C.c_void_p), os_.ctypes.data_as(C.c_void_p),
                                   op.ctypes.data_as(C.c_void_p), oc.ctypes.data_as(C.c_void_p), N.MEM_HOST, device, None))
    return oi, os_, op, oc


def cosine_topk(e: np.ndarray, c: np.ndarray, k: int, device: int = 0):
    """select_top_chunks of the experiment scripts (/root/reference/experiments/retriever/step3_mul.py:255-289,
    continuous_retrieval.py:154-167): cosine similarity then argsort()[-k:][::-1] (ties: higher index first).
    Returns (ids [n_e,k], sims [n_e,k])."""
    e = np.ascontiguousarray(e, dtype=np.float32)
    c = np.ascontiguousarray(c, dtype=np.float32)
    oi = np.full((e.shape[0], k), -1, np.int64)
    os_ = np.full((e.shape[0], k), -_FLT_MAX, np.float32)
    N.check(N.load().vfi_cosine_topk(e.ctypes.data_as(C.c_void_p), e.shape[0], c.ctypes.data_as(C.c_void_p), c.shape[0],
                                     e.shape[1], int(k), os_.ctypes.data_as(C.c_void_p), oi.ctypes.data_as(C.c_void_p),
                                     N.MEM_HOST, device, None))
    return oi, os_
