"""Is the gap between `e2e` and `value` of a pipelined run the copies, or the order in which bench.py measures them (a GPU loses a few
per cent over the first second under load)?  Alternates the two loops of bench.run_dense inside one process.
    python tools/exp_e2e_vs_value.py --workload c3s8"""
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
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--rounds", type=int, default=3)
    args = ap.parse_args()
    ctx = bench.Ctx()
    ctx.rank, ctx.world, ctx.local_rank, ctx.dev = 0, 1, 0, torch.device("cuda", 0)
    torch.cuda.set_device(0)
    w = dict(bench.WORKLOADS[args.workload])
    index, lo, hi = bench.build_dense_index(ctx, w["n"], w["d"], bench.SEED)
    index.set_option(N.OPT_PROFILE, 1)
    q_dev = synth.dense_queries_torch(w["b"], w["d"], bench.SEED, ctx.dev)
    n_buf = 3
    q_pin = [torch.empty((w["b"], w["d"]), dtype=torch.float32).pin_memory() for _ in range(n_buf)]
    for qp in q_pin:
        qp.copy_(q_dev.cpu())
    out_i = [torch.empty((w["b"], w["k"]), dtype=torch.int64).pin_memory() for _ in range(n_buf)]
    out_s = [torch.empty((w["b"], w["k"]), dtype=torch.float32).pin_memory() for _ in range(n_buf)]
    q_stage = [torch.empty((w["b"], w["d"]), dtype=torch.float32, device=ctx.dev) for _ in range(n_buf)]
    h2d, d2h = torch.cuda.Stream(device=ctx.dev), torch.cuda.Stream(device=ctx.dev)
    done = [torch.cuda.Event() for _ in range(n_buf)]

    def value_loop(n):
        prev = None
        for _ in range(n):
            t = index.search_begin(q_dev, w["k"])
            if prev is not None:
                index.search_finish(prev)
            prev = t
        index.search_finish(prev)

    def e2e_loop(n):
        main_s = torch.cuda.current_stream(ctx.dev)
        h2d.wait_stream(main_s)
        d2h.wait_stream(main_s)

        def drain(p):
            ids, scores = index.search_finish(p[0])
            with torch.cuda.stream(d2h):
                out_i[p[1]].copy_(ids, non_blocking=True)
                out_s[p[1]].copy_(scores, non_blocking=True)
                ids.record_stream(d2h)
                scores.record_stream(d2h)
        prev = None
        for i in range(n):
            s = i % n_buf
            with torch.cuda.stream(h2d):
                q_stage[s].copy_(q_pin[s], non_blocking=True)
                done[s].record(h2d)
            main_s.wait_event(done[s])
            t = index.search_begin(q_stage[s], w["k"])
            if prev is not None:
                drain(prev)
            prev = (t, s)
        drain(prev)
        main_s.wait_stream(h2d)
        main_s.wait_stream(d2h)
        torch.cuda.synchronize()

    value_loop(10)
    e2e_loop(3)
    for r in range(args.rounds):
        for name, fn in (("e2e", e2e_loop), ("value", value_loop)):
            index.stats(reset=True)
            ms = bench.timed(ctx, fn, args.steps, host_bound=True)
            st = index.stats()
            print(json.dumps({"workload": args.workload, "loop": name, "round": r, "ms_per_step": round(ms / args.steps, 4),
                              "k1_ms": round(st.fused_ms_total / max(1, st.fused_ms_samples), 4)}), flush=True)


if __name__ == "__main__":
    main()
