"""Multi-GPU parity check, launched one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tools/multi_gpu_check.py
Every rank holds a row shard of one seeded corpus; the fused peer-memory exchange and the NCCL all-gather + merge kernel results must equal the
unsharded CPU oracle bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import flat_ip  # noqa: E402
from veritasfi_b200 import _native as N, synth  # noqa: E402
from veritasfi_b200.dense import DenseIndex  # noqa: E402
from veritasfi_b200.sharded import make_sharded_dense, shard_bounds  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for (n, d, nq, k, path) in [(120_001, 256, 200, 100, N.PATH_FUSED), (50_000, 768, 1, 10, N.PATH_GEMV), (3_000, 64, 9, 20, 0)]:
        xb = synth.dense_corpus_np(n, d, 99)
        xb[n - 1] = xb[7]                      # duplicate on the last shard: cross-shard tie
        xq = synth.dense_queries_np(nq, d, 99, xb)
        xq[0] = xb[7]
        lo, hi = shard_bounds(n, world, rank)
        idx = DenseIndex(d, store="bf16", device=dev, id_offset=lo)
        idx.add(xb[lo:hi])
        idx.set_option(N.OPT_FORCE_PATH, path)
        idx.set_option(N.OPT_TAU_HINT, 1)
        D0, I0 = flat_ip.search(xq, xb, k)
        good = True
        for mode in ("peer", "nccl"):          # fused peer-memory kernel and the all-gather route: same bits
            searcher = make_sharded_dense(idx, exchange=mode)
            for rep in range(3):               # repeated epochs exercise both window parities
                ids, scores = searcher.search(torch.from_numpy(xq).to(dev), k)
                torch.cuda.synchronize()
                good &= bool((ids.cpu().numpy() == I0).all() and (scores.cpu().numpy() == D0).all())
            if searcher.exchange is not None:
                searcher.exchange.close()
        flag = torch.tensor([1 if good else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"[multi-gpu G={world} n={n} d={d} nq={nq} k={k}] all ranks equal to unsharded oracle: {bool(flag.item())}", flush=True)
        ok &= bool(flag.item())
        idx.close()
    # the pipelined sharded searcher (exchange enqueued behind the local search, agreement on re-exchange through the
    # exchange itself) and the sharded hybrid retriever (chunk rows, title rows, BM25 doc ranges; one packed exchange;
    # fusion after the merge) against the CPU oracle, on both exchange routes
    import bench
    ctx = bench.Ctx()
    ctx.rank, ctx.world, ctx.local_rank, ctx.dev = rank, world, local, dev
    for mode in ("peer", "nccl"):
        for fn in (bench.parity_small_dense, bench.parity_small_hybrid):
            res = fn(ctx, mode)
            if rank == 0:
                print(f"[multi-gpu G={world} {fn.__name__} exchange={mode}] {res['what']}: {bool(res['ok'])}", flush=True)
            ok &= bool(res["ok"])
    # the agreement path: rank 0 alone admits nothing (VFI_OPT_TAU_HINT = 2), so every certificate of ITS shard fails; the fail
    # bit travels with the exchange, rank 0 repairs with the exact streaming scorer and ALL ranks exchange the batch again
    for mode in ("peer", "nccl"):
        n, d, nq, k = 90_000, 128, 70, 50
        xb = synth.dense_corpus_np(n, d, 123)
        xq = synth.dense_queries_np(nq, d, 123, xb)
        lo, hi = shard_bounds(n, world, rank)
        idx = DenseIndex(d, store="bf16", device=dev, id_offset=lo)
        idx.add(xb[lo:hi])
        idx.set_option(N.OPT_FORCE_PATH, N.PATH_FUSED)
        idx.set_option(N.OPT_TAU_HINT, 2 if rank == 0 else 1)
        D0, I0 = flat_ip.search(xq, xb, k)
        s = make_sharded_dense(idx, exchange=mode, max_nq=nq, max_k=k)
        q = torch.from_numpy(xq).to(dev)
        t1, t2 = s.search_begin(q, k), s.search_begin(q, k)
        good = True
        for t in (t1, t2):
            ids, scores = s.search_finish(t)
            torch.cuda.synchronize()
            good &= bool((ids.cpu().numpy() == I0).all() and (scores.cpu().numpy() == D0).all())
        good &= s.re_exchanges == 2
        flag = torch.tensor([1 if good else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"[multi-gpu G={world} repair on one rank, exchange={mode}] every rank re-exchanged and equals the oracle: {bool(flag.item())}", flush=True)
        ok &= bool(flag.item())
        if s.exchange is not None:
            s.exchange.close()
        idx.close()
    # the fused push with ranks on different local paths: the last rank scores its shard with the exact streaming scorer, whose
    # rows are pushed by the standalone push kernel, the others send from their rescoring kernel; one epoch on every rank
    if True:
        n, d, nq, k = 90_000, 128, 70, 50
        xb = synth.dense_corpus_np(n, d, 321)
        xq = synth.dense_queries_np(nq, d, 321, xb)
        lo, hi = shard_bounds(n, world, rank)
        idx = DenseIndex(d, store="bf16", device=dev, id_offset=lo)
        idx.add(xb[lo:hi])
        idx.set_option(N.OPT_FORCE_PATH, N.PATH_EXACT if rank == world - 1 else N.PATH_FUSED)
        D0, I0 = flat_ip.search(xq, xb, k)
        s = make_sharded_dense(idx, exchange="peer", max_nq=nq, max_k=k)
        q = torch.from_numpy(xq).to(dev)
        tickets = [s.search_begin(q, k) for _ in range(3)]
        good = True
        for t in tickets:
            ids, scores = s.search_finish(t)
            torch.cuda.synchronize()
            good &= bool((ids.cpu().numpy() == I0).all() and (scores.cpu().numpy() == D0).all())
        good &= s.re_exchanges == 0 and all(t.pushed for t in tickets) == (s.exchange is not None)
        flag = torch.tensor([1 if good else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"[multi-gpu G={world} fused push, last rank on the exact path] three batches in flight equal the oracle: {bool(flag.item())}", flush=True)
        ok &= bool(flag.item())
        if s.exchange is not None:
            s.exchange.close()
        idx.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
