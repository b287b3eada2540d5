"""A/B of one index option inside one process (same box, alternating A B A B so clock drift cancels), pipelined loop.

    python tools/exp_ab.py --workload c3s8 --opt 11 --values 0,1 --rounds 3"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from veritasfi_b200 import _native as N, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3s8")
    ap.add_argument("--opt", type=int, required=True)
    ap.add_argument("--values", default="0,1")
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--steps", type=int, default=50)
    args = ap.parse_args()
    ctx = bench.Ctx()
    ctx.rank, ctx.world, ctx.local_rank, ctx.dev = 0, 1, 0, torch.device("cuda", 0)
    torch.cuda.set_device(0)
    w = dict(bench.WORKLOADS[args.workload])
    index, lo, hi = bench.build_dense_index(ctx, w["n"], w["d"], bench.SEED)
    index.set_option(N.OPT_PROFILE, 1)
    q = synth.dense_queries_torch(w["b"], w["d"], bench.SEED, ctx.dev)
    ref = index.search_batch(q, w["k"])

    def loop(n):
        prev = None
        for _ in range(n):
            t = index.search_begin(q, w["k"])
            if prev is not None:
                index.search_finish(prev)
            prev = t
        return index.search_finish(prev)

    loop(10)
    for r in range(args.rounds):
        for v in (int(x) for x in args.values.split(",")):
            index.set_option(args.opt, v)
            loop(5)
            index.stats(reset=True)
            ms = bench.timed(ctx, loop, args.steps, host_bound=True)
            st = index.stats()
            out = loop(1)
            torch.cuda.synchronize()
            same = bool((out[0] == ref[0]).all() and (out[1] == ref[1]).all())
            print(json.dumps({"workload": args.workload, "opt": args.opt, "value": v, "round": r, "ms_per_step": round(ms / args.steps, 4),
                              "k1_ms": round(st.fused_ms_total / max(1, st.fused_ms_samples), 4),
                              "tail_ms": round(st.tail_ms_total / max(1, st.tail_ms_samples), 4), "same": same}), flush=True)


if __name__ == "__main__":
    main()
