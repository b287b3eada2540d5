import os, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from veritasfi_b200 import _native as N, synth
ctx = bench.Ctx(); ctx.rank, ctx.world, ctx.local_rank, ctx.dev = 0, 1, 0, torch.device("cuda", 0)
torch.cuda.set_device(0)
w = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"])
index, lo, hi = bench.build_dense_index(ctx, w["n"], w["d"], bench.SEED)
index.set_option(N.OPT_PROFILE, 1)
q = synth.dense_queries_torch(w["b"], w["d"], bench.SEED, ctx.dev)
qp = torch.empty((w["b"], w["d"]), dtype=torch.float32).pin_memory(); qp.copy_(q.cpu())
oi = torch.empty((w["b"], w["k"]), dtype=torch.int64).pin_memory(); os_ = torch.empty((w["b"], w["k"]), dtype=torch.float32).pin_memory()
for _ in range(30):
    index.search_host_into(qp.data_ptr(), w["b"], w["k"], os_.data_ptr(), oi.data_ptr(), None)
print("--- steady", file=sys.stderr, flush=True)
ts = []
for _ in range(12):
    t0 = time.perf_counter()
    index.search_host_into(qp.data_ptr(), w["b"], w["k"], os_.data_ptr(), oi.data_ptr(), None)
    ts.append((time.perf_counter() - t0) * 1e3)
print("--- python-observed call ms", [round(t, 3) for t in ts], file=sys.stderr, flush=True)
