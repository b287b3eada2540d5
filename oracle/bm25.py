"""Oracle for the sparse path: bm25s index construction and retrieve as the reference calls them
(/root/reference/src/utils/bm25Retriever.py:14-18,67,75-79) [upstream semantics restated from the
published algorithm].  TEST INFRASTRUCTURE ONLY — see oracle/vfi_oracle.c."""
from __future__ import annotations

import numpy as np

from ._clib import lib, ptr


def build_index(docs, n_vocab: int, k1: float = 1.5, b: float = 0.75):
    """docs: list of token-id lists.  Lucene-variant BM25 impacts precomputed per (token, doc) and
    stored token-major with doc ids ascending (bm25s.BM25.index, method="lucene"):
        idf = ln(1 + (N - df + 0.5)/(df + 0.5));  impact = idf * tf / (tf + k1*(1 - b + b*dl/avgdl))
    Written as plain loops on purpose (independent of the vectorised product builder)."""
    n_docs = len(docs)
    dl = [len(d) for d in docs]
    avgdl = (sum(dl) / n_docs) if n_docs else 0.0
    postings: list[dict[int, int]] = [dict() for _ in range(n_vocab)]
    for di, doc in enumerate(docs):
        for t in doc:
            postings[t][di] = postings[t].get(di, 0) + 1
    indptr = np.zeros(n_vocab + 1, dtype=np.int64)
    indices, data = [], []
    for t in range(n_vocab):
        df = len(postings[t])
        idf = np.log(1.0 + (n_docs - df + 0.5) / (df + 0.5))
        for di in sorted(postings[t]):
            tf = float(postings[t][di])
            tfc = tf / (tf + k1 * (1.0 - b + b * dl[di] / (avgdl if avgdl > 0 else 1.0)))
            indices.append(di)
            data.append(np.float32(idf * tfc))
        indptr[t + 1] = len(indices)
    return indptr, np.asarray(indices, dtype=np.int32), np.asarray(data, dtype=np.float32)


def scores(indptr, indices, data, tokens, n_docs: int) -> np.ndarray:
    """All doc scores of one query: zeros(N) then np.add.at per query token in order (fp32)."""
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    data = np.ascontiguousarray(data, dtype=np.float32)
    tokens = np.ascontiguousarray(tokens, dtype=np.int32)
    out = np.zeros(n_docs, dtype=np.float32)
    lib().vfo_bm25_scores(ptr(indptr), ptr(indices), ptr(data), len(indptr) - 1, ptr(tokens), len(tokens), n_docs, ptr(out))
    return out


def scores_numpy(indptr, indices, data, tokens, n_docs: int) -> np.ndarray:
    """The same thing literally as bm25s' numpy backend writes it (np.add.at) — cross-check and the
    timed CPU baseline of bench.py."""
    out = np.zeros(n_docs, dtype=np.float32)
    nv = len(indptr) - 1
    for t in tokens:
        if 0 <= t < nv:
            s, e = indptr[t], indptr[t + 1]
            np.add.at(out, indices[s:e], data[s:e])
    return out


def retrieve(indptr, indices, data, token_lists, n_docs: int, k: int, id_base: int = 0):
    """Top-k per query under (score desc, id asc).  Returns (ids [B,k], scores [B,k])."""
    from .flat_ip import topk

    B = len(token_lists)
    I = np.empty((B, k), np.int64)
    S = np.empty((B, k), np.float32)
    for q, toks in enumerate(token_lists):
        s = scores(indptr, indices, data, toks, n_docs)
        S[q], I[q] = topk(s, k, id_base)
    return I, S
