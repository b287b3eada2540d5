"""Oracle for the dense path: faiss.normalize_L2 + faiss.IndexFlatIP.search as the reference calls
them (/root/reference/src/utils/faissRetriever.py:18-24,34-37), restated with the canonical score
(sequential fp64 sum over j, rounded once to fp32) and the fixed total order (score desc, id asc).
TEST INFRASTRUCTURE ONLY — see oracle/vfi_oracle.c."""
from __future__ import annotations

import numpy as np

from ._clib import lib, ptr

FLT_MAX = np.finfo(np.float32).max


def _f32c(x) -> np.ndarray:
    return np.ascontiguousarray(x, dtype=np.float32)


def bf16_round(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 (round to nearest even) -> fp32; returns a new array."""
    y = _f32c(x).copy()
    lib().vfo_bf16_round(ptr(y), y.size)
    return y


def normalize_l2(x: np.ndarray) -> np.ndarray:
    """faiss.normalize_L2 semantics, returns a new array (faissRetriever.py:22,35)."""
    y = _f32c(x).copy()
    if y.size:
        lib().vfo_normalize_l2(ptr(y), y.shape[0], y.shape[1])
    return y


def canon_scores(q: np.ndarray, xb: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """Canonical scores of rows `ids` of xb against one query."""
    q, xb = _f32c(q), _f32c(xb)
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    out = np.empty(len(ids), dtype=np.float32)
    lib().vfo_rescore(ptr(q), ptr(xb), xb.shape[1], ptr(ids), len(ids), ptr(out))
    return out


def topk(scores: np.ndarray, k: int, id_base: int = 0):
    scores = _f32c(scores)
    os_, oi = np.empty(k, np.float32), np.empty(k, np.int64)
    lib().vfo_topk(ptr(scores), len(scores), k, id_base, ptr(os_), ptr(oi))
    return os_, oi


def topk_pairs(scores: np.ndarray, ids: np.ndarray, k: int):
    scores = _f32c(scores).ravel()
    ids = np.ascontiguousarray(ids, dtype=np.int64).ravel()
    os_, oi = np.empty(k, np.float32), np.empty(k, np.int64)
    lib().vfo_topk_pairs(ptr(scores), ptr(ids), len(scores), k, ptr(os_), ptr(oi))
    return os_, oi


def search_exhaustive(xq, xb, k: int, id_base: int = 0):
    """IndexFlatIP.search restated exhaustively: O(nq*n*d) canonical dots.  Returns (D, I)."""
    xq, xb = _f32c(xq), _f32c(xb)
    nq, n = xq.shape[0], xb.shape[0]
    d = xq.shape[1]
    D, I = np.empty((nq, k), np.float32), np.empty((nq, k), np.int64)
    lib().vfo_flat_search(ptr(xq), nq, ptr(xb), n, d, k, id_base, ptr(D), ptr(I))
    return D, I


def search(xq, xb, k: int, id_base: int = 0, block: int = 65536):
    """The same result as search_exhaustive at sizes where that is too slow: an fp32 sgemm (torch/MKL,
    blocked over the corpus) proposes 4k candidates per query, the canonical rescoring decides, and
    a certificate (k-th exact score > best excluded sgemm score + eps) proves no excluded row could
    enter; a query that fails it is redone exhaustively."""
    import torch

    xq, xb = _f32c(xq), _f32c(xb)
    nq, n, d = xq.shape[0], xb.shape[0], xq.shape[1]
    kc = min(n, max(4 * k, k + 64))
    if n <= kc or n * nq * d < 5e7:
        return search_exhaustive(xq, xb, k, id_base)
    tq = torch.from_numpy(xq)
    best_s = torch.full((nq, 0), -np.inf)
    best_i = torch.zeros((nq, 0), dtype=torch.int64)
    for r0 in range(0, n, block):
        tb = torch.from_numpy(xb[r0:r0 + block])
        s = tq @ tb.T
        kk = min(kc, s.shape[1])
        ts, ti = torch.topk(s, kk, dim=1)
        best_s = torch.cat([best_s, ts], dim=1)
        best_i = torch.cat([best_i, ti + r0], dim=1)
        if best_s.shape[1] > kc:
            ts, sel = torch.topk(best_s, kc, dim=1)
            best_s, best_i = ts, torch.gather(best_i, 1, sel)
    cand_s, cand_i = best_s.numpy(), best_i.numpy()
    qn = np.linalg.norm(xq.astype(np.float64), axis=1)
    xn = float(np.sqrt((xb.astype(np.float64) ** 2).sum(1).max()))
    D, I = np.empty((nq, k), np.float32), np.empty((nq, k), np.int64)
    for q in range(nq):
        ex = canon_scores(xq[q], xb, cand_i[q])
        ds, di = topk_pairs(ex, cand_i[q] + id_base, k)
        bound = float(cand_s[q].min())
        eps = 4.0 * d * 2.0 ** -24 * qn[q] * xn
        if not (ds[k - 1] > bound + eps):
            ds, di = search_exhaustive(xq[q:q + 1], xb, k, id_base)
            ds, di = ds[0], di[0]
        D[q], I[q] = ds, di
    return D, I


def search_faiss_like(xq, xb, k: int, block: int = 1024, threads: int | None = None):
    """FAISS-behavioural CPU path, kept only to TIME the reference's algorithm (bench.py): fp32 sgemm
    over 1024-row database blocks with a running top-k, the structure of IndexFlatIP.search for
    nq >= 20 [upstream].  Scores carry BLAS summation order, so ids are not the parity reference."""
    import torch

    if threads:
        torch.set_num_threads(threads)
    tq = torch.from_numpy(_f32c(xq))
    tb = torch.from_numpy(_f32c(xb))
    nq, n = tq.shape[0], tb.shape[0]
    kk = min(k, n)
    run_s = torch.full((nq, kk), -float(FLT_MAX))
    run_i = torch.full((nq, kk), -1, dtype=torch.int64)
    big = block * 64  # sgemm over 64 blocks at a time, selection per 1024-row block folded by topk
    for r0 in range(0, n, big):
        s = tq @ tb[r0:r0 + big].T
        ts, ti = torch.topk(s, min(kk, s.shape[1]), dim=1)
        cs = torch.cat([run_s, ts], dim=1)
        ci = torch.cat([run_i, ti + r0], dim=1)
        run_s, sel = torch.topk(cs, kk, dim=1)
        run_i = torch.gather(ci, 1, sel)
    D = np.full((nq, k), -FLT_MAX, np.float32)
    I = np.full((nq, k), -1, np.int64)
    D[:, :kk], I[:, :kk] = run_s.numpy(), run_i.numpy()
    return D, I


def search_blocks(xq, blocks, k: int, kc: int | None = None):
    """`search` for a corpus that does not fit host memory at once (BASELINE configs 3 and 5): `blocks` yields
    (first_row, rows fp32 [m,d]) in any order.  Per block an fp32 sgemm proposes kc candidates per query, which are rescored
    canonically while the block is in memory; every row the block did NOT propose has an sgemm score <= the block's kc-th.
    The result is the exact top-k iff the k-th exact score beats every block's bound + eps; a query that fails the test
    raises (callers pick kc generously: this is a checker, not a product path)."""
    import torch

    xq = _f32c(xq)
    nq, d = xq.shape
    kc = kc or max(4 * k, k + 64)
    tq = torch.from_numpy(xq)
    qn = np.linalg.norm(xq.astype(np.float64), axis=1)
    cand_s = [np.empty(0, np.float32) for _ in range(nq)]
    cand_i = [np.empty(0, np.int64) for _ in range(nq)]
    bound = np.full(nq, -np.inf)
    xn2 = 0.0
    for first, xb in blocks:
        xb = _f32c(xb)
        m = xb.shape[0]
        xn2 = max(xn2, float((xb.astype(np.float64) ** 2).sum(1).max()))
        s = tq @ torch.from_numpy(xb).T
        kk = min(kc, m)
        ts, ti = torch.topk(s, kk, dim=1)
        ts, ti = ts.numpy(), ti.numpy()
        for q in range(nq):
            ex = canon_scores(xq[q], xb, ti[q])
            cand_s[q] = np.concatenate([cand_s[q], ex])
            cand_i[q] = np.concatenate([cand_i[q], ti[q] + first])
            if kk < m:
                bound[q] = max(bound[q], float(ts[q].min()))
            if len(cand_s[q]) > 4 * kc:           # keep the running lists short
                keep_s, keep_i = topk_pairs(cand_s[q], cand_i[q], kc)
                cand_s[q], cand_i[q] = keep_s, keep_i
    D, I = np.empty((nq, k), np.float32), np.empty((nq, k), np.int64)
    xn = float(np.sqrt(xn2))
    for q in range(nq):
        n_have = len(cand_s[q])
        ds, di = topk_pairs(cand_s[q], cand_i[q], min(k, n_have))
        if n_have < k:
            ds = np.concatenate([ds, np.full(k - n_have, -FLT_MAX, np.float32)])
            di = np.concatenate([di, np.full(k - n_have, -1, np.int64)])
        eps = 4.0 * d * 2.0 ** -24 * qn[q] * xn
        if np.isfinite(bound[q]) and not (ds[k - 1] > bound[q] + eps):
            raise RuntimeError(f"search_blocks: query {q} not certified with kc={kc}; raise kc")
        D[q], I[q] = ds, di
    return D, I
