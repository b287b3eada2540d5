"""Corpus row-sharding over the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink).  Rank r holds rows
[r*ceil(N/G), (r+1)*ceil(N/G)) and reports GLOBAL ids; every query batch is replicated.  The only
exchange is one all-gather of the packed per-rank result [B,k] x (fp32 score, int64 id), followed by
the merge kernel (K3) on every rank.  Scores are the canonical rescored values, so the sharded result
is bit-identical to the single-GPU result for any G.

The reference has no counterpart (workers are replicas: /root/reference/experiments/retriever/step3_mul.py:405-446).
On CPU (gloo) the class is exercised with an injected local searcher; the product path needs CUDA."""
from __future__ import annotations

from typing import Callable

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    per = -(-n // world)
    return min(n, rank * per), min(n, (rank + 1) * per)


def pack(scores: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """[B,k] fp32 + [B,k] int64 -> one int64 buffer [B, k + ceil(k/2)] (scores bit-cast pairwise)."""
    B, k = scores.shape
    kpad = k + (k & 1)
    s = torch.zeros((B, kpad), dtype=torch.float32, device=scores.device)
    s[:, :k] = scores
    return torch.cat([ids, s.view(torch.int64)], dim=1).contiguous()


def unpack(buf: torch.Tensor, k: int):
    ids = buf[..., :k].contiguous()
    scores = buf[..., k:].contiguous().view(torch.float32)[..., :k].contiguous()
    return scores, ids


class ShardedSearcher:
    """all-gather + merge around any local searcher `local(q, k) -> (ids [B,k], scores [B,k])`.

    merge: callable (scores [G,B,k], ids [G,B,k], k) -> (ids [B,k], scores [B,k]); the CUDA merge kernel in
    production (veritasfi_b200.dense.merge_topk)."""

    def __init__(self, local: Callable, merge: Callable, group=None):
        self.local = local
        self.merge = merge
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def search(self, q: torch.Tensor, k: int):
        ids, scores = self.local(q, k)
        if self.world == 1:
            return ids, scores
        mine = pack(scores, ids)
        flat = torch.empty((self.world * mine.shape[0], mine.shape[1]), dtype=mine.dtype, device=mine.device)
        dist.all_gather_into_tensor(flat, mine, group=self.group)     # rank-major concatenation along dim 0
        g_scores, g_ids = unpack(flat.view(self.world, mine.shape[0], mine.shape[1]), k)
        return self.merge(g_scores, g_ids, k)


def make_sharded_dense(index, group=None) -> ShardedSearcher:
    """Production wiring: local = DenseIndex.search_batch on this rank's shard, merge = K3 on the GPU."""
    from .dense import merge_topk

    return ShardedSearcher(lambda q, k: index.search_batch(q, k),
                           lambda s, i, k: merge_topk(s, i, k), group)
