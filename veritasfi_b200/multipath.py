"""Fused batch entry for the multi-path retrieval step (SURVEY.md §8b "fused entry", BASELINE configs[3]):
dense chunk path + dense title-summary/table/figure path + BM25 path + rank fusion, all on the GPU, on one GPU or
row/doc-sharded over the GPUs of one box.

    multipath_batch(q_text, q_ts, tokens, k, fusion) -> (ids, scores, path_tag)

Sharding: rank r holds a row shard of the chunk corpus, a row shard of the title corpus and the
doc-range shard of the postings (impacts carry the global idf/avgdl, so doc shards are independent); every shard
reports GLOBAL ids.  The three per-rank lists [P, B, L] travel in ONE exchange (the fused peer-memory push + merge
kernel, or one NCCL all-gather + the merge kernel, over P*B rows); the title -> chunk map, the de-duplication, the
dropping of BM25's zero-score filler and the rank fusion run after the merge in one kernel (vfi_fuse_hybrid).

Path priority for the union is the reference's: dense > title-summary > BM25
(/root/reference/src/utils/ensembleRetriever.py:62,137,187)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as N
from . import fusion as F
from .bm25_compat import GpuPostings
from .dense import DenseIndex

_FLT_MAX = float(np.finfo(np.float32).max)


def fuse_hybrid(ids: torch.Tensor, scores: torch.Tensor, title_to_chunk: torch.Tensor | None, title_path: int, sparse_path: int,
                k: int, k_rrf: float = 60.0, layout: str = "pbl"):
    """ids int64 / scores float32, [P,B,L] (layout "pbl", what the exchange produces) or [B,P,L] ("bpl"), cuda.
    Returns (ids [B,k], fused scores [B,k])."""
    ids, scores = ids.contiguous(), scores.contiguous()
    if layout == "pbl":
        P, B, L = ids.shape
        path_stride, query_stride = B * L, L
    else:
        B, P, L = ids.shape
        path_stride, query_stride = L, P * L
    dev = ids.device
    oi = torch.empty((B, k), dtype=torch.int64, device=dev)
    os_ = torch.empty((B, k), dtype=torch.float32, device=dev)
    t2c = C.c_void_p(title_to_chunk.data_ptr()) if title_to_chunk is not None else None
    n_t = int(title_to_chunk.numel()) if title_to_chunk is not None else 0
    N.check(N.load().vfi_fuse_hybrid(C.c_void_p(ids.data_ptr()), C.c_void_p(scores.data_ptr()), B, P, L, path_stride, query_stride,
                                     t2c, n_t, title_path if title_to_chunk is not None else -1, sparse_path, float(k_rrf), int(k),
                                     C.c_void_p(os_.data_ptr()), C.c_void_p(oi.data_ptr()), dev.index or 0,
                                     C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return oi, os_


class MultiPathRetriever:
    """chunks / titles: this rank's DenseIndex shards (id offsets set, so they report global rows); title_to_chunk: int64
    [N_ts_total] on the GPU, GLOBAL title row -> GLOBAL chunk row; postings: this rank's doc-range GpuPostings.
    sharded: a veritasfi_b200.sharded.RowExchange (or None on one GPU)."""

    def __init__(self, chunks: DenseIndex, titles: DenseIndex | None, title_to_chunk: torch.Tensor | None,
                 postings: GpuPostings | None, depth: int = 200, sharded=None):
        self.chunks = chunks
        self.titles = titles
        self.title_to_chunk = title_to_chunk
        self.postings = postings
        self.depth = depth
        self.sharded = sharded
        self.paths = ["chunks"] + (["titles"] if titles is not None else []) + (["bm25"] if postings is not None else [])
        self.title_path = 1 if titles is not None else -1
        self.sparse_path = (len(self.paths) - 1) if postings is not None else -1

    def path_lists(self, q_text: torch.Tensor, q_ts: torch.Tensor | None, tokens):
        """The per-path ranked lists of this rank, path-major: (ids int64 [P,B,L], scores float32 [P,B,L]) with global ids
        (title path: global TITLE rows).  The two dense searches are enqueued back to back; their certificates are read
        after the BM25 kernels have been enqueued behind them."""
        dev = q_text.device
        B, L, P = q_text.shape[0], self.depth, len(self.paths)
        ids = torch.empty((P, B, L), dtype=torch.int64, device=dev)
        scores = torch.empty((P, B, L), dtype=torch.float32, device=dev)
        tickets = [(self.chunks, self.chunks.search_begin(q_text, L, out=(ids[0], scores[0])))]
        if self.titles is not None:
            tickets.append((self.titles, self.titles.search_begin(q_text if q_ts is None else q_ts, L, out=(ids[1], scores[1]))))
        if self.postings is not None:
            # tokens: per-query id lists, or the packed (toks int32, qptr int64) pair of GpuPostings.pack_tokens
            toks, qptr = tokens if isinstance(tokens, tuple) else GpuPostings.pack_tokens(tokens)
            p = self.sparse_path
            self.postings.search_csr_device(toks, qptr, L, dev, out=(ids[p], scores[p]))   # results stay in HBM
        for index, t in tickets:
            index.search_finish(t)
        return ids, scores

    def multipath_batch(self, q_text: torch.Tensor, q_ts: torch.Tensor | None, tokens, k: int, fusion: str = "rrf",
                        k_rrf: float = 60.0):
        ids, scores = self.path_lists(q_text, q_ts, tokens)
        P, B, L = ids.shape
        if self.sharded is not None and self.sharded.world > 1:
            mi, ms = self.sharded.exchange_rows(scores.view(P * B, L), ids.view(P * B, L), L)   # ONE exchange for all paths
            ids, scores = mi.view(P, B, L), ms.view(P, B, L)
        if fusion == "rrf":
            fi, fs = fuse_hybrid(ids, scores, self.title_to_chunk, self.title_path, self.sparse_path, k, k_rrf, "pbl")
            return fi, fs, None
        if fusion == "union":
            # the reference's ordered de-dup union at the id level (not a benchmark path: tensor glue is fine here)
            lists_i, lists_s = [ids[0]], [scores[0]]
            if self.titles is not None:
                it, st = ids[1], scores[1]
                mapped = torch.where(it >= 0, self.title_to_chunk[it.clamp_min(0)], it)
                di, ds, _, _ = F.union(mapped.view(B, 1, L).contiguous(), st.view(B, 1, L).contiguous())
                lists_i.append(di)
                lists_s.append(ds)
            if self.postings is not None:
                lists_i.append(ids[self.sparse_path])
                lists_s.append(scores[self.sparse_path])
            ui, us, up, _ = F.union(torch.stack(lists_i, dim=1).contiguous(), torch.stack(lists_s, dim=1).contiguous())
            return ui[:, :k], us[:, :k], up[:, :k]
        raise ValueError("fusion must be 'rrf' or 'union'")
