"""Seeded synthetic inputs of the BASELINE.json shapes.  numpy for host-side test
cases, torch generators for device-side benchmark corpora (generated per shard with seed + rank)."""
from __future__ import annotations

import numpy as np


def bf16_round_np(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 (round to nearest even) -> fp32 with numpy integer arithmetic."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) & 0xFFFF0000).astype(np.uint32).view(np.float32).reshape(x.shape)


def dense_corpus_np(n: int, d: int, seed: int, dup_frac: float = 0.001, bf16: bool = True):
    """Rows N(0,1) -> L2-normalised -> (optionally) rounded to bf16; a fraction of rows are exact
    duplicates of earlier rows so that the tie-break is exercised."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    if bf16:
        x = bf16_round_np(x)
    n_dup = int(n * dup_frac)
    if n_dup and n > 1:
        dst = rng.choice(np.arange(1, n), size=n_dup, replace=False)
        src = (rng.random(n_dup) * dst).astype(np.int64)
        x[dst] = x[src]
    return np.ascontiguousarray(x)


def dense_queries_np(b: int, d: int, seed: int, corpus: np.ndarray | None = None, near_frac: float = 0.25, bf16: bool = True):
    """Queries by the same recipe; a fraction are noisy copies of corpus rows (true near neighbours)."""
    rng = np.random.default_rng(seed + 7919)
    q = rng.standard_normal((b, d), dtype=np.float32)
    if corpus is not None and len(corpus):
        n_near = int(b * near_frac)
        rows = rng.integers(0, len(corpus), size=n_near)
        q[:n_near] = corpus[rows] + 0.05 * q[:n_near]
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    if bf16:
        q = bf16_round_np(q)
    return np.ascontiguousarray(q)


def dense_corpus_torch(n: int, d: int, seed: int, device, dtype="bf16", chunk: int = 1 << 18, dup_frac: float = 0.001):
    """Device-side corpus shard for the benchmark: same recipe, generated in chunks on the GPU."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n, d), dtype=torch.bfloat16 if dtype == "bf16" else torch.float32, device=device)
    for r0 in range(0, n, chunk):
        r1 = min(n, r0 + chunk)
        x = torch.randn((r1 - r0, d), generator=g, device=device, dtype=torch.float32)
        x = torch.nn.functional.normalize(x, dim=1)
        out[r0:r1] = x.to(out.dtype)
    n_dup = int(n * dup_frac)
    if n_dup and n > 1:
        dst = torch.randint(1, n, (n_dup,), generator=g, device=device)
        src = (torch.rand(n_dup, generator=g, device=device) * dst).long()
        out[dst] = out[src]
    return out


def dense_queries_torch(b: int, d: int, seed: int, device, bf16: bool = True):
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed + 7919)
    q = torch.nn.functional.normalize(torch.randn((b, d), generator=g, device=device, dtype=torch.float32), dim=1)
    if bf16:
        q = q.to(torch.bfloat16).to(torch.float32)
    return q.contiguous()


def zipf_postings(n_docs: int, n_vocab: int, seed: int, mean_len: int = 128, s: float = 1.1):
    """Synthetic token postings: Zipf(s) term popularity, doc length ~ Poisson(mean_len).
    Returns (doc_ptr int64 [N+1], doc_tokens int64 [total]) — a CSR of token ids per document."""
    rng = np.random.default_rng(seed)
    ranks = np.arange(1, n_vocab + 1, dtype=np.float64)
    p = ranks ** (-s)
    p /= p.sum()
    cdf = np.cumsum(p)
    dl = np.maximum(1, rng.poisson(mean_len, size=n_docs)).astype(np.int64)
    doc_ptr = np.zeros(n_docs + 1, dtype=np.int64)
    np.cumsum(dl, out=doc_ptr[1:])
    u = rng.random(int(doc_ptr[-1]))
    toks = np.searchsorted(cdf, u).astype(np.int64)
    np.minimum(toks, n_vocab - 1, out=toks)
    return doc_ptr, toks


def bm25_queries(b: int, n_vocab: int, seed: int, tmin: int = 4, tmax: int = 16, s: float = 1.1):
    """Queries of U{tmin..tmax} tokens, half drawn from the Zipf law and half uniformly."""
    rng = np.random.default_rng(seed + 104729)
    ranks = np.arange(1, n_vocab + 1, dtype=np.float64)
    p = ranks ** (-s)
    cdf = np.cumsum(p / p.sum())
    out = []
    for _ in range(b):
        t = int(rng.integers(tmin, tmax + 1))
        z = np.minimum(np.searchsorted(cdf, rng.random(t)), n_vocab - 1)
        uni = rng.integers(0, n_vocab, size=t)
        pick = rng.random(t) < 0.5
        out.append(np.where(pick, z, uni).astype(np.int32).tolist())
    return out


def zipf_postings_torch(n_docs: int, n_vocab: int, seed: int, device, mean_len: int = 128, s: float = 1.1, doc_chunk: int = 1 << 20):
    """The same synthetic postings generated on the GPU for one doc shard (BASELINE configs[3] is 5 M docs, 6.4e8 tokens:
    too slow to build with numpy).  Returns token-major COUNTS, not impacts: (tok int64 [nnz], doc int64 [nnz] ascending per
    token, tf int64 [nnz], dl int64 [n_docs]) — impacts need the GLOBAL df and avgdl (`bm25_impacts_torch`)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    ranks = torch.arange(1, n_vocab + 1, dtype=torch.float64, device=device)
    cdf = torch.cumsum(ranks.pow(-s) / ranks.pow(-s).sum(), dim=0)
    dl_all, keys = [], []
    for d0 in range(0, n_docs, doc_chunk):
        nd = min(doc_chunk, n_docs - d0)
        dl = torch.poisson(torch.full((nd,), float(mean_len), device=device), generator=g).clamp_min(1).long()
        total = int(dl.sum())
        u = torch.rand(total, generator=g, device=device, dtype=torch.float64)
        tok = torch.searchsorted(cdf, u).clamp_max(n_vocab - 1)
        doc = torch.repeat_interleave(torch.arange(d0, d0 + nd, device=device), dl)
        k = torch.unique(tok * n_docs + doc, return_counts=True)     # (token, doc) pairs of this chunk with their tf
        keys.append(k)
        dl_all.append(dl)
        del u, tok, doc
    key = torch.cat([k[0] for k in keys])
    tf = torch.cat([k[1] for k in keys])
    del keys
    key, order = torch.sort(key)                                       # token-major, doc ascending
    tf = tf[order]
    del order
    tok = key // n_docs
    doc = key - tok * n_docs
    return tok, doc, tf, torch.cat(dl_all)


def bm25_impacts_torch(tok, doc, tf, dl, n_vocab: int, n_docs_total: int, df_global, avgdl: float, k1: float = 1.5, b: float = 0.75):
    """bm25s-lucene impacts (the formula of bm25_compat.build_csc) for one doc shard on the GPU, from GLOBAL document
    frequencies and average length.  Returns (indptr int64 [V+1], indices int32 [nnz], data float32 [nnz]) cuda tensors."""
    import torch

    df = df_global.double()
    idf = torch.log(1.0 + (n_docs_total - df + 0.5) / (df + 0.5))
    tfd = tf.double()
    tfc = tfd / (tfd + k1 * (1.0 - b + b * dl[doc].double() / (avgdl if avgdl > 0 else 1.0)))
    data = (idf[tok] * tfc).float()
    indptr = torch.zeros(n_vocab + 1, dtype=torch.int64, device=tok.device)
    indptr[1:] = torch.cumsum(torch.bincount(tok, minlength=n_vocab), dim=0)
    return indptr, doc.int(), data
