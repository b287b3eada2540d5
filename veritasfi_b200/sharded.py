"""Corpus row-sharding over the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink).  Rank r holds rows
[r*ceil(N/G), (r+1)*ceil(N/G)) and reports GLOBAL ids; every query batch is replicated.  The only
exchange is one all-gather of the packed per-rank result [B,k] x (fp32 score, int64 id), followed by
the merge kernel (K3) on every rank.  Scores are the canonical rescored values, so the sharded result
is bit-identical to the single-GPU result for any G.

On NVLink boxes the exchange and the merge are ONE kernel per rank over peer memory (`PeerExchange`,
csrc/peer_exchange.cuh): every rank stores its results straight into windows in its peers' HBM, signals per-query
flags and merges what it received — no pack/unpack kernels and no NCCL launch on the data path (NCCL is only used
once, to hand round the 64-byte IPC handles).  The all-gather route stays as the portable path (gloo on CPU, or
VFI_EXCHANGE=nccl).

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


class PeerExchangeUnavailable(RuntimeError):
    """Raised on EVERY rank when any rank could not set up its peer windows."""


class PeerExchange:
    """Receive windows in every rank's HBM, mapped by all peers through CUDA IPC (include/vfi.h: vfi_exchange_*).

    `merge(scores, ids, k_out)` is a collective: every rank calls it with its own [B,k] shard result and gets the
    global top-k_out, bit-identical on all ranks.  Global ids must be < 2^32 - 1."""

    def __init__(self, device: torch.device, max_nq: int, max_k: int, group=None):
        import ctypes as C

        from . import _native as N
        self._N, self._C = N, C
        self.device = device
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.max_nq, self.max_k = int(max_nq), int(max_k)
        self._h = C.c_void_p()
        lib = N.load()
        # Local steps may fail on one rank only (no IPC between these processes, out of memory ...); the collectives
        # below are executed by every rank regardless, and the verdict is shared, so the ranks never diverge.
        err = None
        mine = (C.c_uint8 * N.IPC_HANDLE_BYTES)()
        try:
            N.check(lib.vfi_exchange_create(device.index or 0, self.rank, self.world, self.max_nq, self.max_k, C.byref(self._h)))
            N.check(lib.vfi_exchange_handle(self._h, mine))
        except Exception as e:   # noqa: BLE001
            err = e
        # the handles travel once over whatever backend the group has (NCCL here); the data path never uses it
        t = torch.tensor(list(bytes(mine)), dtype=torch.uint8, device=device)
        allh = torch.empty(self.world * N.IPC_HANDLE_BYTES, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(allh, t, group=group)
        if err is None:
            try:
                buf = (C.c_uint8 * (self.world * N.IPC_HANDLE_BYTES)).from_buffer_copy(bytes(allh.cpu().numpy().tobytes()))
                N.check(lib.vfi_exchange_connect(self._h, buf))
            except Exception as e:   # noqa: BLE001
                err = e
        bad = torch.tensor([0 if err is None else 1], dtype=torch.int32, device=device)
        dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=group)   # also the barrier: every window is zeroed and mapped
        if int(bad.item()):
            self.close(barrier=False)     # the all-reduce above already ordered every rank; nobody has pushed yet
            raise PeerExchangeUnavailable(str(err) if err is not None else "setup failed on another rank")

    def merge(self, scores: torch.Tensor, ids: torch.Tensor, k_out: int):
        C, N = self._C, self._N
        B, k = scores.shape
        scores, ids = scores.contiguous(), ids.contiguous()
        out_s = torch.empty((B, k_out), dtype=torch.float32, device=self.device)
        out_i = torch.empty((B, k_out), dtype=torch.int64, device=self.device)
        N.check(N.load().vfi_exchange_merge(self._h, C.c_void_p(scores.data_ptr()), C.c_void_p(ids.data_ptr()), B, k, int(k_out),
                                            C.c_void_p(out_s.data_ptr()), C.c_void_p(out_i.data_ptr()),
                                            C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return out_i, out_s

    def close(self, barrier: bool = True) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            torch.cuda.synchronize(self.device)
            if barrier and dist.is_initialized():
                dist.barrier(group=self.group)   # no peer is still storing into a window that is about to be freed
            self._N.load().vfi_exchange_destroy(self._h)
            self._h = self._C.c_void_p()


class ShardedSearcher:
    """exchange + merge around any local searcher `local(q, k) -> (ids [B,k], scores [B,k])`.

    merge: callable (scores [G,B,k], ids [G,B,k], k) -> (ids [B,k], scores [B,k]); the CUDA merge kernel in
    production (veritasfi_b200.dense.merge_topk).  exchange: a PeerExchange (fused push + merge over NVLink peer
    memory) or None (all-gather over the process group, then `merge`)."""

    def __init__(self, local: Callable, merge: Callable, group=None, exchange: "PeerExchange | None" = None):
        self.local = local
        self.merge = merge
        self.group = group
        self.exchange = exchange
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def search(self, q: torch.Tensor, k: int):
        ids, scores = self.local(q, k)
        if self.world == 1:
            return ids, scores
        if self.exchange is not None and scores.shape[0] <= self.exchange.max_nq and k <= self.exchange.max_k:
            return self.exchange.merge(scores, ids, k)
        mine = pack(scores, ids)
        flat = torch.empty((self.world * mine.shape[0], mine.shape[1]), dtype=mine.dtype, device=mine.device)
        dist.all_gather_into_tensor(flat, mine, group=self.group)     # rank-major concatenation along dim 0
        g_scores, g_ids = unpack(flat.view(self.world, mine.shape[0], mine.shape[1]), k)
        return self.merge(g_scores, g_ids, k)


def make_sharded_dense(index, group=None, exchange: str | None = None, max_nq: int = 1024, max_k: int = 256) -> ShardedSearcher:
    """Production wiring: local = DenseIndex.search_batch on this rank's shard, merge = K3 on the GPU.

    exchange: "peer" (fused push + merge kernel over NVLink peer memory), "nccl" (all-gather + merge kernel), or None =
    the VFI_EXCHANGE environment variable, default "peer"."""
    import os

    from .dense import merge_topk

    mode = (exchange or os.environ.get("VFI_EXCHANGE", "peer")).lower()
    if mode not in ("peer", "nccl"):
        raise ValueError("exchange must be 'peer' or 'nccl'")
    ex = None
    if mode == "peer" and dist.is_initialized() and dist.get_world_size(group) > 1:
        # peer windows need CUDA IPC between the ranks' processes; if any rank cannot map its peers every rank falls
        # back to the all-gather route together (PeerExchange shares the verdict, so the ranks never disagree)
        try:
            ex = PeerExchange(index.device, max_nq, max_k, group)
        except PeerExchangeUnavailable as e:
            import warnings
            warnings.warn(f"peer-memory exchange unavailable ({e}); using the NCCL all-gather route")
    return ShardedSearcher(lambda q, k: index.search_batch(q, k),
                           lambda s, i, k: merge_topk(s, i, k), group, ex)
