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

    def merge(self, scores: torch.Tensor, ids: torch.Tensor, k_out: int, fail_ptrs=(None, None), any_fail: torch.Tensor | None = None):
        """fail_ptrs: up to two device addresses of certificate counters (DenseIndex.ticket_flag_ptr) of the searches that
        produced these rows; any_fail: int32 [1] cuda tensor (zeroed by the caller) that every rank finds set to 1 when any
        rank's rows were not final — see vfi_exchange_merge_flagged."""
        C, N = self._C, self._N
        B, k = scores.shape
        scores, ids = scores.contiguous(), ids.contiguous()
        out_s = torch.empty((B, k_out), dtype=torch.float32, device=self.device)
        out_i = torch.empty((B, k_out), dtype=torch.int64, device=self.device)
        fa, fb = (list(fail_ptrs) + [None, None])[:2]
        N.check(N.load().vfi_exchange_merge_flagged(
            self._h, C.c_void_p(scores.data_ptr()), C.c_void_p(ids.data_ptr()), B, k, int(k_out), C.c_void_p(out_s.data_ptr()),
            C.c_void_p(out_i.data_ptr()), C.c_void_p(fa) if fa else None, C.c_void_p(fb) if fb else None,
            C.c_void_p(any_fail.data_ptr()) if any_fail is not None else None,
            C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return out_i, out_s

    def merge_pushed(self, B: int, k: int, k_out: int, fail_ptr: int | None = None):
        """The wait + merge half for rows that a search begun with `DenseIndex.search_begin(..., push=self)` is pushing (or has
        pushed) on the current stream; fail_ptr: the batch's certificate counter (DenseIndex.ticket_flag_ptr).  Returns (ids [B,k_out], scores [B,k_out], slot); `any_fail(slot)` is valid once an event
        recorded behind this call has completed.  A collective, like merge."""
        C, N = self._C, self._N
        out_s = torch.empty((B, k_out), dtype=torch.float32, device=self.device)
        out_i = torch.empty((B, k_out), dtype=torch.int64, device=self.device)
        slot = C.c_int(-1)
        N.check(N.load().vfi_exchange_merge_pushed(self._h, B, int(k), int(k_out), C.c_void_p(out_s.data_ptr()),
                                                   C.c_void_p(out_i.data_ptr()), C.c_void_p(fail_ptr) if fail_ptr else None,
                                                   C.byref(slot),
                                                   C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return out_i, out_s, slot.value

    def any_fail(self, slot: int) -> bool:
        v = self._C.c_int(0)
        self._N.check(self._N.load().vfi_exchange_any_fail(self._h, int(slot), self._C.byref(v)))
        return v.value != 0

    def close(self, barrier: bool = True) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            torch.cuda.synchronize(self.device)
            if barrier and dist.is_initialized():
                dist.barrier(group=self.group)   # no peer is still storing into a window that is about to be freed
            self._N.load().vfi_exchange_destroy(self._h)
            self._h = self._C.c_void_p()


class ShardTicket:
    """One sharded batch in flight (ShardedSearcher.search_begin)."""
    __slots__ = ("local", "out", "slot", "event", "k", "pushed")

    def __init__(self, local, out, slot, event, k, pushed=False):
        self.local, self.out, self.slot, self.event, self.k, self.pushed = local, out, slot, event, k, pushed


class ShardedSearcher:
    """exchange + merge around any local searcher `local(q, k) -> (ids [B,k], scores [B,k])`.

    merge: callable (scores [G,B,k], ids [G,B,k], k) -> (ids [B,k], scores [B,k]); the CUDA merge kernel in
    production (veritasfi_b200.dense.merge_topk).  exchange: a PeerExchange (fused push + merge over NVLink peer
    memory) or None (all-gather over the process group, then `merge`).  index: the DenseIndex behind `local` — enables
    the pipelined form search_begin / search_finish (CUDA only)."""

    def __init__(self, local: Callable, merge: Callable, group=None, exchange: "PeerExchange | None" = None, index=None):
        self.local = local
        self.merge = merge
        self.group = group
        self.exchange = exchange
        self.index = index
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._flags_host = None     # pinned ring of "any rank repaired" flags, one per batch in flight
        self._flags_dev = None
        self._next_slot = 0
        self.re_exchanges = 0
        import os
        self.fused_push = os.environ.get("VFI_FUSED_PUSH", "1") != "0"   # 0: the unfused exchange kernel behind the search

    def exchange_rows(self, scores: torch.Tensor, ids: torch.Tensor, k_out: int, fail_ptrs=(None, None), any_fail=None):
        """Global top-k_out of every row over the ranks: scores/ids [R,k] per rank (global ids) -> (ids, scores) [R,k_out],
        identical on all ranks.  Rows are independent, so the three lists of the hybrid retriever travel as R = 3*B rows."""
        R, k = scores.shape
        if self.world == 1:
            return ids, scores
        if self.exchange is not None and R <= self.exchange.max_nq and k <= self.exchange.max_k:
            return self.exchange.merge(scores, ids, k_out, fail_ptrs, any_fail)
        mine = pack(scores, ids)
        flat = torch.empty((self.world * mine.shape[0], mine.shape[1]), dtype=mine.dtype, device=mine.device)
        dist.all_gather_into_tensor(flat, mine, group=self.group)     # rank-major concatenation along dim 0
        g_scores, g_ids = unpack(flat.view(self.world, mine.shape[0], mine.shape[1]), k)
        if any_fail is not None:      # the all-gather route: the ranks agree through one more (tiny) collective
            mine_fail = torch.zeros(1, dtype=torch.int32, device=scores.device)
            for p in fail_ptrs:
                if p:
                    mine_fail = torch.maximum(mine_fail, _int32_at(p, scores.device).clamp(max=1))
            dist.all_reduce(mine_fail, op=dist.ReduceOp.MAX, group=self.group)
            any_fail.copy_(mine_fail)
        return self.merge(g_scores, g_ids, k_out)

    def search(self, q: torch.Tensor, k: int):
        ids, scores = self.local(q, k)
        return self.exchange_rows(scores, ids, k)

    # -- pipelined form: batch i+1 is enqueued (local search AND exchange) before the host looks at batch i ------------
    def search_begin(self, q: torch.Tensor, k: int) -> ShardTicket:
        """Enqueue the local search and, right behind it on the stream, the exchange of its (not yet certified) rows.  The
        exchange carries this rank's certificate counter to every peer; search_finish repairs what failed and, if ANY rank
        had to, all ranks exchange that batch once more — they agree on it without a host collective."""
        if self.index is None:
            raise RuntimeError("search_begin needs the DenseIndex behind the local searcher (make_sharded_dense)")
        if (self.world > 1 and self.exchange is not None and self.fused_push and q.shape[0] <= self.exchange.max_nq
                and k <= self.exchange.max_k):
            # compute + collective: the rescoring kernel of the local search sends every finished row to the peers itself (with
            # that query's certificate verdict), the kernel behind it waits for the peers' rows and merges; the "any rank has
            # to repair" bit reaches the host through mapped memory — no pack, copy or memset operation in the stream
            t = self.index.search_begin(q, k, push=self.exchange)
            oi, os_, slot = self.exchange.merge_pushed(q.shape[0], k, k, self.index.ticket_flag_ptr(t))
            ev = None
            if q.is_cuda:            # (a CPU stand-in for index and exchange exercises this control flow under gloo)
                ev = torch.cuda.Event()
                ev.record()
            return ShardTicket(t, (oi, os_), slot, ev, k, True)
        t = self.index.search_begin(q, k)
        if self.world == 1:
            return ShardTicket(t, None, -1, None, k)
        dev = q.device
        if self._flags_host is None:
            self._flags_host = torch.zeros(64, dtype=torch.int32).pin_memory()
            self._flags_dev = torch.zeros(64, dtype=torch.int32, device=dev)
        slot = self._next_slot
        self._next_slot = (slot + 1) % 64
        any_fail = self._flags_dev[slot:slot + 1]
        any_fail.zero_()
        out = self.exchange_rows(t.scores, t.ids, k, (self.index.ticket_flag_ptr(t), None), any_fail)
        self._flags_host[slot:slot + 1].copy_(any_fail, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return ShardTicket(t, out, slot, ev, k)

    def search_finish(self, ticket: ShardTicket):
        ids, scores = self.index.search_finish(ticket.local)     # waits for the local batch, repairs flagged queries
        if self.world == 1:
            return ids, scores
        if ticket.event is not None:
            ticket.event.synchronize()
        failed = self.exchange.any_fail(ticket.slot) if ticket.pushed else int(self._flags_host[ticket.slot]) != 0
        if failed:                                                # the same value on every rank
            self.re_exchanges += 1
            return self.exchange_rows(scores, ids, ticket.k)
        return ticket.out


def _int32_at(ptr: int, device) -> torch.Tensor:
    """A [1] int32 cuda tensor aliasing a device address owned by libvfi (a certificate counter)."""
    import ctypes as C

    class _Arr:
        __cuda_array_interface__ = {"shape": (1,), "typestr": "<i4", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(_Arr(), device=device)


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
                           lambda s, i, k: merge_topk(s, i, k), group, ex, index=index)


def make_row_exchange(device, exchange: str | None = None, max_rows: int = 3072, max_k: int = 256, group=None) -> ShardedSearcher:
    """The exchange half alone (`exchange_rows`) for callers that produce their per-rank rows themselves — the hybrid
    retriever sends the lists of its three paths as 3*B rows in one exchange."""
    import os

    from .dense import merge_topk

    mode = (exchange or os.environ.get("VFI_EXCHANGE", "peer")).lower()
    if mode not in ("peer", "nccl"):
        raise ValueError("exchange must be 'peer' or 'nccl'")
    ex = None
    if mode == "peer" and dist.is_initialized() and dist.get_world_size(group) > 1:
        try:
            ex = PeerExchange(torch.device(device), max_rows, max_k, group)
        except PeerExchangeUnavailable as e:
            import warnings
            warnings.warn(f"peer-memory exchange unavailable ({e}); using the NCCL all-gather route")
    return ShardedSearcher(None, lambda s, i, k: merge_topk(s, i, k), group, ex)
