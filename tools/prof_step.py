"""A bare loop of hot-path steps for ncu (no parity checks, no CPU arm, no e2e): build the workload of bench.py and run
`--steps` steps after `--warmup`.

    python tools/prof_step.py --workload c2 --steps 3 --warmup 2
    python tools/prof_step.py --workload online --steps 3        # IndexFlatIP.search(k = 2048), 1M x 1024 fp32, nq = 4
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from veritasfi_b200 import _native as N, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--tau-m", type=int, default=0)
    args = ap.parse_args()
    ctx = bench.Ctx()
    ctx.rank, ctx.world, ctx.local_rank, ctx.dev = 0, 1, 0, torch.device("cuda", 0)
    torch.cuda.set_device(0)
    if args.workload == "online":
        from veritasfi_b200.dense import DenseIndex
        idx = DenseIndex(1024, store="f32", device=ctx.dev)
        for c in range(4):
            idx.add(synth.dense_corpus_torch(250_000, 1024, 5 + c, ctx.dev, dtype="f32"))
        q = synth.dense_queries_torch(4, 1024, 5, ctx.dev, bf16=False)
        for _ in range(args.warmup + args.steps):
            idx.search_batch(q, 2048)
        torch.cuda.synchronize()
        return
    w = dict(bench.WORKLOADS[args.workload])
    if w["kind"] == "hybrid":
        from veritasfi_b200.multipath import MultiPathRetriever
        chunks, titles, t2c, postings, tokens, info = bench.build_hybrid(ctx, w, bench.SEED)
        for i in (chunks, titles):
            i.set_option(N.OPT_TAU_M, args.tau_m)
        mp = MultiPathRetriever(chunks, titles, t2c, postings, depth=w["depth"])
        q = synth.dense_queries_torch(w["b"], w["d"], bench.SEED, ctx.dev)
        for _ in range(args.warmup + args.steps):
            mp.multipath_batch(q, None, tokens, w["k"])
        torch.cuda.synchronize()
        return
    index, lo, hi = bench.build_dense_index(ctx, w["n"], w["d"], bench.SEED)
    index.set_option(N.OPT_TAU_M, args.tau_m)
    index.set_option(N.OPT_PROFILE, 0)
    q = synth.dense_queries_torch(w["b"], w["d"], bench.SEED, ctx.dev)
    for _ in range(args.warmup + args.steps):
        index.search_batch(q, w["k"])
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
