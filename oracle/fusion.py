"""Oracle for rank fusion.  TEST INFRASTRUCTURE ONLY — see oracle/vfi_oracle.c.

union: the reference's fusion — one shared `seen_ids` set across the FAISS, Title-Summary and BM25
sections (/root/reference/src/utils/ensembleRetriever.py:58,72-74,148-150,194-196), at the id level.
rrf: reciprocal-rank fusion (north_star; absent from the reference, SURVEY.md finding 4)."""
from __future__ import annotations

import numpy as np

from ._clib import lib, ptr


def rrf(ids: np.ndarray, k_rrf: float, k: int):
    """ids int64 [B,P,L] (-1 padding) -> (ids [B,k], scores [B,k])."""
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    B, P, L = ids.shape
    oi, os_ = np.empty((B, k), np.int64), np.empty((B, k), np.float32)
    for b in range(B):
        lib().vfo_rrf(ptr(ids[b]), P, L, float(k_rrf), k, ptr(os_[b]), ptr(oi[b]))
    return oi, os_


def rrf_python(ids: np.ndarray, k_rrf: float, k: int):
    """The textbook formula in plain Python/numpy float32 (cross-check of the C restatement)."""
    B, P, L = ids.shape
    oi = np.full((B, k), -1, np.int64)
    os_ = np.full((B, k), -np.finfo(np.float32).max, np.float32)
    for b in range(B):
        acc: dict[int, np.float32] = {}
        for p in range(P):
            for r in range(L):
                d = int(ids[b, p, r])
                if d < 0:
                    continue
                term = np.float32(1.0) / (np.float32(k_rrf) + np.float32(r + 1))
                acc[d] = np.float32(acc.get(d, np.float32(0.0)) + term)
        order = sorted(acc.items(), key=lambda kv: (-float(kv[1]), kv[0]))[:k]
        for j, (d, s) in enumerate(order):
            oi[b, j], os_[b, j] = d, s
    return oi, os_


def union(ids: np.ndarray, scores: np.ndarray):
    """ids int64 [B,P,L], scores float32 [B,P,L] -> (ids [B,P*L], scores, path int32, count int32[B])."""
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    B, P, L = ids.shape
    oi = np.empty((B, P * L), np.int64)
    os_ = np.empty((B, P * L), np.float32)
    op = np.empty((B, P * L), np.int32)
    cnt = np.empty(B, np.int32)
    for b in range(B):
        cnt[b] = lib().vfo_union(ptr(ids[b]), ptr(scores[b]), P, L, ptr(oi[b]), ptr(os_[b]), ptr(op[b]))
    return oi, os_, op, cnt


def union_python(ids: np.ndarray, scores: np.ndarray):
    """The reference's loop shape: a seen set walked in path order then rank order."""
    B, P, L = ids.shape
    out = []
    for b in range(B):
        seen, row = set(), []
        for p in range(P):
            for r in range(L):
                d = int(ids[b, p, r])
                if d < 0 or d in seen:
                    continue
                seen.add(d)
                row.append((d, float(scores[b, p, r]), p))
        out.append(row)
    return out


def hybrid(ids: np.ndarray, scores: np.ndarray, title_to_chunk, title_path: int, sparse_path: int, k_rrf: float, k: int):
    """The hybrid fusion step, stage by stage (ids/scores [B,P,L]): title ids -> chunk ids with the first occurrence kept
    and the ranks closed up (union over that one path), BM25 entries with score <= 0 dropped (bm25s' zero-score filler),
    then RRF.  Restates what vfi_fuse_hybrid does in one launch."""
    ids = np.array(ids, dtype=np.int64, copy=True)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    if title_path >= 0:
        t = ids[:, title_path, :]
        mapped = np.where(t >= 0, np.asarray(title_to_chunk, dtype=np.int64)[np.clip(t, 0, None)], -1)
        mi, _, _, _ = union(mapped[:, None, :], scores[:, title_path, :][:, None, :])
        ids[:, title_path, :] = mi
    if sparse_path >= 0:
        ids[:, sparse_path, :] = np.where(scores[:, sparse_path, :] > 0, ids[:, sparse_path, :], -1)
    return rrf(ids, k_rrf, k)
