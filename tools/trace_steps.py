"""VFI_TRACE_STEPS=1 python tools/trace_steps.py [--workload c3s8] [--profile 0|1]: the pipelined loop of bench.py for a few
batches with one event behind every stream operation; the library prints the per-operation device times on stderr."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from veritasfi_b200 import _native as N, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3s8")
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--profile", type=int, default=1)
    args = ap.parse_args()
    ctx = bench.Ctx()
    ctx.rank, ctx.world, ctx.local_rank, ctx.dev = 0, 1, 0, torch.device("cuda", 0)
    torch.cuda.set_device(0)
    w = dict(bench.WORKLOADS[args.workload])
    index, lo, hi = bench.build_dense_index(ctx, w["n"], w["d"], bench.SEED)
    index.set_option(N.OPT_PROFILE, args.profile)
    q = synth.dense_queries_torch(w["b"], w["d"], bench.SEED, ctx.dev)

    def loop(n):
        prev = None
        for _ in range(n):
            t = index.search_begin(q, w["k"])
            if prev is not None:
                index.search_finish(prev)
            prev = t
        index.search_finish(prev)

    loop(20)
    torch.cuda.synchronize()
    print(f"--- {args.workload} profile={args.profile}", file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    loop(args.steps)
    torch.cuda.synchronize()
    print(f"--- wall {1e3 * (time.perf_counter() - t0) / args.steps:.3f} ms per step", file=sys.stderr, flush=True)


if __name__ == "__main__":
    main()
