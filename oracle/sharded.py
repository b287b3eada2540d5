"""Oracle for the multi-GPU path: virtual ranks.  Shard the corpus row-wise, search every shard with
global ids, concatenate (the all-gather) and merge under the total order.  Must equal the unsharded
oracle for any shard count.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import numpy as np

from . import flat_ip


def shard_bounds(n: int, g: int):
    """Rank r holds rows [r*ceil(n/g), min(n,(r+1)*ceil(n/g)))."""
    per = -(-n // g)
    return [(min(n, r * per), min(n, (r + 1) * per)) for r in range(g)]


def merge(scores: np.ndarray, ids: np.ndarray, k: int):
    """scores/ids [G,B,k_in] -> (ids [B,k], scores [B,k])."""
    G, B, k_in = scores.shape
    oi, os_ = np.empty((B, k), np.int64), np.empty((B, k), np.float32)
    for b in range(B):
        os_[b], oi[b] = flat_ip.topk_pairs(scores[:, b, :].ravel(), ids[:, b, :].ravel(), k)
    return oi, os_


def search_sharded(xq, xb, k: int, g: int, exhaustive: bool = True):
    parts_s, parts_i = [], []
    fn = flat_ip.search_exhaustive if exhaustive else flat_ip.search
    for lo, hi in shard_bounds(len(xb), g):
        D, I = fn(xq, xb[lo:hi], k, id_base=lo)
        parts_s.append(D)
        parts_i.append(I)
    return merge(np.stack(parts_s), np.stack(parts_i), k)
