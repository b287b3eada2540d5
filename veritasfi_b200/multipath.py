"""Fused batch entry for the multi-path retrieval step (SURVEY.md §8b "fused entry"):
dense chunk path + dense title-summary/table/figure path + BM25 path + rank fusion, all on the GPU.

    multipath_batch(q_text, q_ts, tokens, k, fusion) -> (ids, scores, path_tag)

Path priority for the union is the reference's: dense > title-summary > BM25
(/root/reference/src/utils/ensembleRetriever.py:62,137,187)."""
from __future__ import annotations

import numpy as np
import torch

from . import fusion as F
from .bm25_compat import GpuPostings
from .dense import DenseIndex


class MultiPathRetriever:
    def __init__(self, chunks: DenseIndex, titles: DenseIndex | None, title_to_chunk: torch.Tensor | None,
                 postings: GpuPostings | None, depth: int = 200):
        self.chunks = chunks
        self.titles = titles
        self.title_to_chunk = title_to_chunk  # int64 [N_ts] on the GPU: row of the chunk a title vector stands for
        self.postings = postings
        self.depth = depth

    def multipath_batch(self, q_text: torch.Tensor, q_ts: torch.Tensor | None, tokens, k: int, fusion: str = "rrf",
                        k_rrf: float = 60.0):
        dev = q_text.device
        B, L = q_text.shape[0], self.depth
        lists_i, lists_s = [], []
        i0, s0 = self.chunks.search_batch(q_text, L)
        lists_i.append(i0)
        lists_s.append(s0)
        if self.titles is not None:
            it, st = self.titles.search_batch(q_text if q_ts is None else q_ts, L)
            mapped = torch.where(it >= 0, self.title_to_chunk[it.clamp_min(0)], it)
            # several title vectors can stand for one chunk: keep the first (best-ranked) occurrence
            di, ds, _, _ = F.union(mapped.view(B, 1, L), st.view(B, 1, L))
            lists_i.append(di)
            lists_s.append(ds)
        if self.postings is not None:
            # tokens: per-query id lists, or the packed (toks int32, qptr int64) pair of GpuPostings.pack_tokens
            toks, qptr = tokens if isinstance(tokens, tuple) else GpuPostings.pack_tokens(tokens)
            bi, bs = self.postings.search_csr_device(toks, qptr, L, dev)     # results stay in HBM for the fusion kernel
            lists_i.append(bi)
            lists_s.append(bs)
        ids = torch.stack(lists_i, dim=1).contiguous()
        scores = torch.stack(lists_s, dim=1).contiguous()
        if fusion == "rrf":
            fi, fs = F.rrf(ids, k, k_rrf)
            return fi, fs, None
        if fusion == "union":
            ui, us, up, uc = F.union(ids, scores)
            return ui[:, :k], us[:, :k], up[:, :k]
        raise ValueError("fusion must be 'rrf' or 'union'")
