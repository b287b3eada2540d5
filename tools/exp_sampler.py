import json, os, sys, time
import torch
sys.path.insert(0, "/root/repo")
import bench
from veritasfi_b200 import _native as N, synth
ctx = bench.Ctx(); ctx.rank, ctx.world, ctx.local_rank, ctx.dev = 0, 1, 0, torch.device("cuda", 0)
torch.cuda.set_device(0)
w = dict(bench.WORKLOADS["c3s8"])
index, lo, hi = bench.build_dense_index(ctx, w["n"], w["d"], bench.SEED)
index.set_option(N.OPT_PROFILE, 1)
q = synth.dense_queries_torch(w["b"], w["d"], bench.SEED, ctx.dev)
def loop(n):
    prev = None
    for _ in range(n):
        t = index.search_begin(q, w["k"])
        if prev is not None:
            index.search_finish(prev)
        prev = t
    index.search_finish(prev)
loop(200)
for r in range(4):
    for on in (1, 0):
        s = bench.ClockSampler(0)
        if on:
            s.start(); s.wait_first()
        ms = bench.timed(ctx, loop, 200)
        if on:
            s.stop()
        print(json.dumps({"sampler": on, "round": r, "ms_per_step": round(ms / 200, 4)}), flush=True)
