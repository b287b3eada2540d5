#!/usr/bin/env python
"""bench.py — the retrieval hot path on B200, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4|c5|...] [--impl reference]

Metric (BASELINE.json): queries/sec at 10M x 1024, top-100, 1024-query batch (config C3, the default); the corpus
is row-sharded over the N GPUs of the box (strong scaling: 10M rows in total for every N).
A step = one pass of the hot path over one query batch: prepare queries -> fused tcgen05 GEMM+top-k'
-> candidate reduction -> exact rescoring + certificate -> (N>1: exchange + merge).  Two batches are kept in flight
(batch i+1 is enqueued before the host reads the certificate of batch i), at every N.

  value     whole-job queries/s with queries and corpus resident in HBM
  e2e       the same with HOST buffers: N=1 through the reference-facing C-ABI call (vfi_index_search with pinned host
            queries in, host ids/scores out); N>1 through the sharded searcher with the pinned H2D copy of every
            batch and the D2H read of its results inside the timed region
  roofline  the dominant kernel: algorithmic FLOPs (or bytes) per launch over its CUDA-event duration, against
            MEASURED_PEAKS.json; `rooflines` lists every kernel of a multi-kernel workload (c4)
  parity    what was timed is also checked: a small seeded corpus through the same sharded searcher against the CPU
            oracle, and canonical rescoring + order of rows returned on the timed corpus
  cpu_baseline  the reference's CPU algorithm (oracle port of IndexFlatIP.search: blocked fp32 sgemm + running top-k)
            on the host cores: a bounded row slice per step (scaled), plus one full-corpus step when RAM allows

Workloads: c3 (default, BASELINE configs[2]), c2 (configs[1]), c4 (configs[3]: hybrid 3-path + RRF, 5M chunks),
c5 (configs[4]: single-query latency over 50M x 768, p50/p99), and shard-sized variants for profiling.
--impl reference runs only the CPU arm (rank 0) and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "c3": dict(kind="dense", n=10_000_000, d=1024, b=1024, k=100, desc="10Mx1024 bf16 corpus, 1024-query batch, top-100 (BASELINE configs[2])"),
    "c2": dict(kind="dense", n=1_000_000, d=1024, b=256, k=100, desc="1Mx1024 bf16 corpus, 256-query batch, top-100 (BASELINE configs[1])"),
    "c5": dict(kind="latency", n=50_000_000, d=768, b=1, k=10, desc="50Mx768 bf16 corpus, single query, top-10, p50/p99 (BASELINE configs[4])"),
    "c5s8": dict(kind="latency", n=6_250_000, d=768, b=1, k=10, desc="one 1/8 row shard of C5 (6.25Mx768 bf16), single query, top-10"),
    "c4": dict(kind="hybrid", n=5_000_000, n_ts=1_000_000, d=1024, b=1024, k=50, depth=200, vocab=262_144, mean_len=128,
               desc="hybrid 3-path: dense 5Mx1024 + dense 1Mx1024 title vectors + BM25 (V=262144, ~96 unique terms/doc), depth 200, RRF-60, top-50, 1024-query batch (BASELINE configs[3])"),
    "c4s8": dict(kind="hybrid", n=625_000, n_ts=125_000, d=1024, b=1024, k=50, depth=200, vocab=262_144, mean_len=128,
                 desc="one 1/8 shard of C4 on one GPU (625k chunks, 125k titles)"),
    "c4small": dict(kind="hybrid", n=60_000, n_ts=12_000, d=256, b=64, k=20, depth=50, vocab=8192, mean_len=40, desc="smoke-size hybrid"),
    "c3s8": dict(kind="dense", n=1_250_000, d=1024, b=1024, k=100, desc="one 1/8 row shard of C3 (1.25Mx1024 bf16), 1024-query batch, top-100"),
    "c3q": dict(kind="dense", n=2_500_000, d=1024, b=1024, k=100, desc="a quarter of C3 (2.5Mx1024 bf16; on two GPUs each shard is the 1.25M rows of an 8-GPU C3 run), 1024-query batch, top-100"),
    "d200": dict(kind="dense", n=625_000, d=1024, b=1024, k=200, desc="the chunk path of one C4 shard alone (625kx1024 bf16, 1024 queries, depth 200: k' = 256)"),
    "b16": dict(kind="dense", n=1_000_000, d=1024, b=16, k=100, desc="1Mx1024 bf16 corpus, 16-query batch, top-100 (serving batch)"),
    "b64": dict(kind="dense", n=1_000_000, d=1024, b=64, k=100, desc="1Mx1024 bf16 corpus, 64-query batch, top-100 (serving batch)"),
    "b128": dict(kind="dense", n=1_000_000, d=1024, b=128, k=100, desc="1Mx1024 bf16 corpus, 128-query batch, top-100 (serving batch)"),
    "small": dict(kind="dense", n=200_000, d=1024, b=256, k=100, desc="200kx1024 smoke-size workload"),
}
SEED = 1234 + 2


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]),
                    tf_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def wait_first(self, timeout=2.0):
        t_end = time.time() + timeout
        while self.proc is not None and not self.rows and time.time() < t_end:
            time.sleep(0.01)
        self.mark = len(self.rows)      # samples from here on lie inside (or right at the edge of) the timed region

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        mark = getattr(self, "mark", 0)
        n_end = len(self.rows)
        if n_end <= mark:               # region shorter than one sampling period: take the next sample
            time.sleep(0.08)
            n_end = len(self.rows)
        self.proc.terminate()
        try:                            # the sampler's own teardown (a driver client going away) must not overlap the NEXT timed region
            self.proc.wait(timeout=3)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.rows[mark:n_end] or self.rows[-1:]
        for r in rows:
            parts = [x.strip() for x in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def read_traffic(workload: str):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/ncu_traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)[workload]
        return float(t["dram_bytes_read"] + t["dram_bytes_write"])
    except Exception:
        return None


def base_config(name: str, w: dict) -> dict:
    """The workload description both arms print (identical dicts: the driver compares them)."""
    cfg = {"workload": f"{name}: {w['desc']}", "corpus_rows": w["n"], "dim": w["d"], "batch": w["b"], "k": w["k"],
           "l2": "inputs exceed L2 (corpus shard >> 126 MB); no flush needed"}
    if w["kind"] == "hybrid":
        cfg.update(title_rows=w["n_ts"], depth=w["depth"], vocab=w["vocab"], fusion="rrf-60")
    return cfg


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_dense_sample(w, rows: int, steps: int, warmup: int, threads: int | None):
    """The reference's CPU retrieval path (IndexFlatIP.search restated: blocked fp32 sgemm + running top-k) on the host
    cores over a contiguous row slice of the same synthetic workload.  Returns (seconds per sampled step, slice arrays)."""
    import torch
    from oracle import flat_ip
    from veritasfi_b200 import synth

    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    xb = synth.dense_corpus_np(rows, w["d"], SEED, dup_frac=0.0, bf16=True)
    xq = synth.dense_queries_np(w["b"], w["d"], SEED, None, bf16=True)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        flat_ip.search_faiss_like(xq, xb, w["k"], threads=cores)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sum(times) / len(times), xb, xq, cores


def cpu_dense_full(w, xb_slice, xq, cores: int):
    """ONE step over a corpus of the full size when host RAM holds it in fp32 (else None): the seeded slice tiled to n rows,
    every tile rolled along the dimension by its index so that no two rows are equal.  sgemm and top-k time do not depend
    on the values; this is the measured counterpart of the scaled per-step figure."""
    import numpy as np
    import psutil
    from oracle import flat_ip

    need = w["n"] * w["d"] * 4
    if psutil.virtual_memory().available < 1.6 * need or need > 64e9:
        return None
    t_build = time.perf_counter()
    xb = np.empty((w["n"], w["d"]), dtype=np.float32)
    rows = len(xb_slice)
    for i, r0 in enumerate(range(0, w["n"], rows)):
        r1 = min(w["n"], r0 + rows)
        xb[r0:r1] = np.roll(xb_slice[: r1 - r0], i, axis=1)
    t_build = time.perf_counter() - t_build
    t0 = time.perf_counter()
    flat_ip.search_faiss_like(xq, xb, w["k"], threads=cores)
    dt = time.perf_counter() - t0
    return {"value": w["b"] / dt, "unit": "queries/s", "ms_per_step": dt * 1e3, "steps": 1, "rows": w["n"],
            "corpus": f"slice tiled to {w['n']} rows (rolled per tile), built in {t_build:.1f} s"}


def cpu_hybrid_sample(w, rows: int, steps: int, warmup: int):
    """CPU arm of the hybrid workload: both dense paths (sgemm port on a row slice, scaled), BM25 (the oracle's np.add.at
    order restatement + top-k on a doc slice, scaled) and RRF (C oracle) for the full batch.  Seconds per FULL step."""
    import numpy as np
    from oracle import bm25 as obm, fusion as ofu
    from veritasfi_b200 import synth
    from veritasfi_b200.bm25_compat import build_csc

    wd = dict(w)
    t_dense, xb, xq, cores = cpu_dense_sample(wd, rows, steps, warmup, None)
    scale_d = (w["n"] + w["n_ts"]) / rows
    docs = min(w["n"], 100_000)
    doc_ptr, toks = synth.zipf_postings(docs, w["vocab"], SEED, mean_len=w["mean_len"])
    csc = build_csc(doc_ptr, toks, w["vocab"])
    nq_s = min(w["b"], 32)
    qs = synth.bm25_queries(nq_s, w["vocab"], SEED)
    t0 = time.perf_counter()
    bi, bs = obm.retrieve(*csc, qs, docs, w["depth"])
    t_bm = (time.perf_counter() - t0) * (w["b"] / nq_s) * (w["n"] / docs)
    lists = np.stack([bi, bi, bi], axis=1)
    t0 = time.perf_counter()
    ofu.rrf(lists, 60.0, w["k"])
    t_fuse = (time.perf_counter() - t0) * (w["b"] / nq_s)
    t_full = t_dense * scale_d + t_bm + t_fuse
    sample = (f"dense: {rows} of {w['n'] + w['n_ts']} rows x {w['b']} queries (x{scale_d:.1f}); bm25: {docs} of {w['n']} docs x {nq_s} of "
              f"{w['b']} queries (scaled); rrf: {nq_s} queries (scaled); {cores} threads")
    return t_full, t_dense, sample, cores


def reference_arm(args, name, w, config, steps, warmup):
    rows = args.cpu_rows
    # exactly K timed steps after W warm-ups; a step is one pass over a bounded row slice, scaled to the corpus.  Bounded so the
    # whole run ends within minutes whatever K is: at most ~120 s of CPU work for the sampled steps.
    est = 0.3 * (rows / 250_000) * (w["b"] / 1024) * (w["d"] / 1024)
    while rows > 20_000 and est * (steps + warmup) > 120:
        rows //= 2
        est /= 2
    rows = min(rows, w["n"])
    if w["kind"] == "hybrid":
        t_full, t_step, sample, cores = cpu_hybrid_sample(w, rows, steps, warmup)
        scale = t_full / t_step
        full = None
    else:
        t_step, xb, xq, cores = cpu_dense_sample(w, rows, steps, warmup, None)
        scale = w["n"] / rows
        import torch
        sample = (f"{rows} of {w['n']} rows x {w['b']} queries per step, time scaled x{scale:.1f}; torch {torch.__version__} sgemm, "
                  f"{cores} threads")
        full = cpu_dense_full(w, xb, xq, cores) if not args.no_full_cpu_step else None
    value = w["b"] / (t_step * scale)
    cpu = {"value": value, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample, "sample_scale": scale}
    if full is not None:
        cpu["measured_full"] = full
    line = {"impl": "reference", "metric": "queries/sec", "value": value, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "ms_per_step is the measured time of one sampled step; value = batch / (ms_per_step * sample_scale)"}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm helpers
class Ctx:
    pass


def timed(ctx, fn, k_steps, host_bound=False):
    """K steps bracketed by barrier + synchronize on both sides, CUDA events on the launching stream, max over ranks."""
    import torch
    import torch.distributed as dist
    if ctx.world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    fn(k_steps)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = max(e0.elapsed_time(e1), 0.0)
    if host_bound:            # steps that end in a host synchronisation are bounded below by wall time
        ms = max(ms, wall * 1e3)
    t = torch.tensor([ms], dtype=torch.float64, device=ctx.dev)
    if ctx.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def max_over_ranks(ctx, x: float) -> float:
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=ctx.dev)
    if ctx.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_ranks_ok(ctx, ok: bool) -> bool:
    import torch
    import torch.distributed as dist
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=ctx.dev)
    if ctx.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


def build_dense_index(ctx, n_total, d, seed, store="bf16"):
    """This rank's row shard of a seeded synthetic corpus, generated on the device chunk by chunk."""
    import torch
    from veritasfi_b200 import synth
    from veritasfi_b200.dense import DenseIndex
    from veritasfi_b200.sharded import shard_bounds
    lo, hi = shard_bounds(n_total, ctx.world, ctx.rank)
    index = DenseIndex(d, store=store, device=ctx.dev, id_offset=lo)
    index.reserve(hi - lo)
    chunk = 1 << 20
    for r0 in range(0, hi - lo, chunk):
        r1 = min(hi - lo, r0 + chunk)
        index.add(synth.dense_corpus_torch(r1 - r0, d, seed + 1000 * ctx.rank + (r0 // chunk), ctx.dev))
    torch.cuda.synchronize()
    return index, lo, hi


def parity_small_dense(ctx, exchange: str):
    """A small seeded corpus sharded exactly like the timed one, through the same searcher (same exchange mode, pipelined
    form), against the CPU oracle on rank 0; all ranks must hold the oracle's bits."""
    import numpy as np
    import torch
    from oracle import flat_ip
    from veritasfi_b200 import synth
    from veritasfi_b200.dense import DenseIndex
    from veritasfi_b200.sharded import make_sharded_dense, shard_bounds
    n, d, nq, k = 120_001, 256, 96, 100
    xb = synth.dense_corpus_np(n, d, 99)
    xb[n - 1] = xb[7]                      # duplicate on the last shard: cross-shard tie
    xq = synth.dense_queries_np(nq, d, 99, xb)
    xq[0] = xb[7]
    lo, hi = shard_bounds(n, ctx.world, ctx.rank)
    idx = DenseIndex(d, store="bf16", device=ctx.dev, id_offset=lo)
    idx.add(xb[lo:hi])
    s = make_sharded_dense(idx, exchange=exchange, max_nq=nq, max_k=k)
    q = torch.from_numpy(xq).to(ctx.dev)
    D0, I0 = flat_ip.search(xq, xb, k)
    ok = True
    t1 = s.search_begin(q, k)
    t2 = s.search_begin(q, k)
    for t in (t1, t2):
        ids, scores = s.search_finish(t)
        torch.cuda.synchronize()
        ok &= bool((ids.cpu().numpy() == I0).all() and (scores.cpu().numpy() == D0).all())
    ids, scores = s.search(q, k)
    ok &= bool((ids.cpu().numpy() == I0).all() and (scores.cpu().numpy() == D0).all())
    ok = all_ranks_ok(ctx, ok)
    if s.exchange is not None:
        s.exchange.close()
    idx.close()
    return {"ok": ok, "what": f"{n}x{d} seeded corpus, {nq} queries, top-{k}, sharded over {ctx.world} GPU(s), exchange={exchange if ctx.world > 1 else 'none'}: "
                              "ids and scores equal to oracle.flat_ip.search on every rank (pipelined and synchronous form)"}


def parity_timed_rows(ctx, index, lo, hi, q_dev, ids, scores, n_check=4):
    """On the timed corpus: for a few queries, the rows this rank owns among the returned ids are read back and rescored in the
    canonical order by the oracle; the returned scores must be those bits and the rows must be in (score desc, id asc) order."""
    import numpy as np
    from oracle import flat_ip
    ok = True
    checked = 0
    ids_h, sc_h = ids[:n_check].cpu().numpy(), scores[:n_check].cpu().numpy()
    q_h = flat_ip.bf16_round(q_dev[:n_check].cpu().numpy())
    for j in range(len(ids_h)):
        order_ok = all((sc_h[j][i] > sc_h[j][i + 1]) or (sc_h[j][i] == sc_h[j][i + 1] and ids_h[j][i] < ids_h[j][i + 1])
                       for i in range(len(ids_h[j]) - 1))
        ok &= bool(order_ok)
        mine = [(p, int(r)) for p, r in enumerate(ids_h[j]) if lo <= r < hi][:32]
        for p, r in mine:
            row = index.read_rows(r - lo, 1)
            want = flat_ip.canon_scores(q_h[j], row, np.array([0]))[0]
            ok &= bool(want == sc_h[j][p])
            checked += 1
    ok = all_ranks_ok(ctx, ok)
    return {"ok": ok, "what": f"timed corpus: first {n_check} queries, returned rows read back per shard and rescored by the oracle "
                              f"({checked} rows on rank 0), order checked"}


# ------------------------------------------------------------------------------------------------ dense workloads
def run_dense(args, name, w, ctx, config, steps, warmup, result_out):
    import torch
    from veritasfi_b200 import _native as N, synth
    from veritasfi_b200.sharded import make_sharded_dense

    index, lo, hi = build_dense_index(ctx, w["n"], w["d"], SEED)
    n_local = hi - lo
    index.set_option(N.OPT_TAU_HINT, args.hint)
    index.set_option(N.OPT_PROFILE, 1)
    if args.pair is not None:
        index.set_option(N.OPT_CTA_PAIR, args.pair)
    index.set_option(N.OPT_TAU_M, args.tau_m)
    if args.no_small:
        index.set_option(N.OPT_SMALL_BATCH, 1)
    searcher = make_sharded_dense(index, exchange=args.exchange, max_nq=w["b"], max_k=w["k"])
    q_dev = synth.dense_queries_torch(w["b"], w["d"], SEED, ctx.dev)
    n_buf = 3
    q_pin = [torch.empty((w["b"], w["d"]), dtype=torch.float32).pin_memory() for _ in range(n_buf)]
    for qp in q_pin:
        qp.copy_(q_dev.cpu())
    out_i_pin = [torch.empty((w["b"], w["k"]), dtype=torch.int64).pin_memory() for _ in range(n_buf)]
    out_s_pin = [torch.empty((w["b"], w["k"]), dtype=torch.float32).pin_memory() for _ in range(n_buf)]
    q_stage = [torch.empty((w["b"], w["d"]), dtype=torch.float32, device=ctx.dev) for _ in range(n_buf)]
    pipelined = not args.sync

    def run_device_steps(k_steps):
        """K steps of the hot path with inputs resident in HBM.  A serving loop keeps two batches in flight so the GPU does not
        idle while the host looks at the certificate flag; every batch is still certified (and repaired if needed, and
        exchanged again if any rank repaired) inside the timed region."""
        if not pipelined:
            for _ in range(k_steps):
                searcher.search(q_dev, w["k"])
            return
        prev = None
        for _ in range(k_steps):
            t = searcher.search_begin(q_dev, w["k"])
            if prev is not None:
                searcher.search_finish(prev)
            prev = t
        searcher.search_finish(prev)

    e2e_stream = torch.cuda.Stream(device=ctx.dev).cuda_stream if args.e2e_stream == "caller" else None

    def run_e2e_host_call(k_steps):
        for _ in range(k_steps):   # the reference-facing host call of the C ABI: H2D, search, D2H inside
            index.search_host_into(q_pin[0].data_ptr(), w["b"], w["k"], out_s_pin[0].data_ptr(), out_i_pin[0].data_ptr(), e2e_stream)

    h2d_stream = torch.cuda.Stream(device=ctx.dev)
    d2h_stream = torch.cuda.Stream(device=ctx.dev)
    h2d_done = [torch.cuda.Event() for _ in range(n_buf)]

    def run_e2e_pipelined(k_steps):
        """Host buffers through the sharded searcher: per batch a pinned H2D copy of its queries, the search (+ exchange),
        a D2H read of its ids and scores; two batches in flight.  The copies run on streams of their own (H2D of batch i+1 and
        D2H of batch i-1 overlap the kernels of batch i) and every one of them lies inside the timed region.  search_finish
        returns when the batch is complete on the device, so its D2H copy needs no event from the main stream — an event
        recorded there would sit behind the NEXT batch's kernels and hold the copy (and the H2D queued behind it) back."""
        main = torch.cuda.current_stream(ctx.dev)
        h2d_stream.wait_stream(main)
        d2h_stream.wait_stream(main)

        def drain(prev):
            ids, scores = searcher.search_finish(prev[0])
            with torch.cuda.stream(d2h_stream):
                out_i_pin[prev[1]].copy_(ids, non_blocking=True)
                out_s_pin[prev[1]].copy_(scores, non_blocking=True)
                ids.record_stream(d2h_stream)
                scores.record_stream(d2h_stream)

        prev = None
        for i in range(k_steps):
            s = i % n_buf
            with torch.cuda.stream(h2d_stream):
                q_stage[s].copy_(q_pin[s], non_blocking=True)
                h2d_done[s].record(h2d_stream)
            main.wait_event(h2d_done[s])
            t = searcher.search_begin(q_stage[s], w["k"])
            if prev is not None:
                drain(prev)
            prev = (t, s)
        drain(prev)
        main.wait_stream(h2d_stream)
        main.wait_stream(d2h_stream)
        torch.cuda.synchronize()

    run_e2e = run_e2e_host_call if ctx.world == 1 else run_e2e_pipelined

    parity = {"small": parity_small_dense(ctx, args.exchange)}
    run_device_steps(warmup)
    run_e2e(min(warmup, 3))
    index.stats(reset=True)
    sampler = ClockSampler(ctx.local_rank)
    if ctx.rank == 0:
        sampler.start()
        sampler.wait_first()
    launches0 = N.launch_count()
    ms_total = timed(ctx, run_device_steps, steps)
    launches = N.launch_count() - launches0
    st = index.stats()
    clocks = sampler.stop() if ctx.rank == 0 else None
    index.stats(reset=True)
    ms_e2e = timed(ctx, run_e2e, steps, host_bound=True)
    st_e2e = index.stats()
    ms_e2e_pipe = timed(ctx, run_e2e_pipelined, steps, host_bound=True) if ctx.world == 1 else None
    ids, scores = searcher.search(q_dev, w["k"])
    torch.cuda.synchronize()
    parity["timed"] = parity_timed_rows(ctx, index, lo, hi, q_dev, ids, scores)
    kernel_ms = max_over_ranks(ctx, st.fused_ms_total / max(1, st.fused_ms_samples))

    if ctx.rank == 0:
        peaks = read_peaks()
        ms_step = ms_total / steps
        flops_per_launch = 2.0 * w["b"] * n_local * w["d"]
        achieved_tf = flops_per_launch / (kernel_ms * 1e-3) / 1e12 if kernel_ms > 0 else 0.0
        bytes_per_launch = float(n_local) * w["d"] * 2
        gbs = bytes_per_launch / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
        path = int(st.last_path)
        if path == N.PATH_FUSED and w["b"] >= 211:
            roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                        "frac": achieved_tf / peaks["tf_sustained"],
                        "traffic": read_traffic(name) if ctx.world == 1 else None,
                        "kernel": "dense_fused_pair_kernel<MODE_TOPK>", "kernel_ms": kernel_ms,
                        "peak_source": f"{peaks['source']} MEASURED_PEAKS.json bf16_tflops_sustained",
                        "algorithmic": f"2*B*(N/G)*d = {flops_per_launch:.4g} FLOP per launch; (N/G)*d*2 = {bytes_per_launch:.4g} B",
                        "traffic_source": "ncu dram__bytes_read+write per launch, profiles/ncu_traffic.json",
                        "hbm_frac": gbs / peaks["hbm"]}
        else:
            kname = {N.PATH_FUSED: "dense_small_kernel (swapped operands, query block resident)" if 8 < w["b"] <= 64 and not args.no_small
                     else "dense_fused_pair_kernel<MODE_TOPK>", N.PATH_GEMV: "gemv_topk_kernel", N.PATH_EXACT: "exact_scores_kernel"}[path]
            roofline = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                        "traffic": read_traffic(name) if ctx.world == 1 else None, "kernel": kname, "kernel_ms": kernel_ms,
                        "peak_source": f"{peaks['source']} MEASURED_PEAKS.json hbm_gbs",
                        "algorithmic": f"(N/G)*d*2 = {bytes_per_launch:.4g} B per launch (B = {w['b']} < 211 FLOP/B machine balance: HBM-bound)",
                        "tensor_frac": achieved_tf / peaks["tf_sustained"]}
        line = {"metric": "queries/sec", "value": w["b"] / (ms_step * 1e-3), "unit": "queries/s", "n_gpus": ctx.world, "steps": steps,
                "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic", "config": config, "roofline": roofline, "clocks": clocks,
                "e2e": {"value": w["b"] / (ms_e2e / steps * 1e-3), "unit": "queries/s", "ms_per_step": ms_e2e / steps,
                        "h2d_bytes_per_step": w["b"] * w["d"] * 4, "d2h_bytes_per_step": w["b"] * w["k"] * 12,
                        "kernel_ms": st_e2e.fused_ms_total / max(1, st_e2e.fused_ms_samples),
                        "tail_ms": st_e2e.tail_ms_total / max(1, st_e2e.tail_ms_samples),
                        "how": ("vfi_index_search (C ABI) with pinned host queries in, host ids/scores out, one synchronous call per step"
                                if ctx.world == 1 else
                                "sharded searcher, two batches in flight: pinned H2D of the queries, search + exchange, D2H of ids/scores per batch")},
                "gpu_launches": int(launches), "parity": parity,
                "run": {"pipeline": ("two batches in flight: search_begin(i+1) (local search and exchange) is enqueued before search_finish(i) "
                                     "reads the certificate flag of batch i" if pipelined else "one synchronous search (+ exchange) per step"),
                        "sharding": f"row-sharded over {ctx.world} GPU(s)" + (f", exchange={args.exchange}" if ctx.world > 1 else ""),
                        "re_exchanges": int(searcher.re_exchanges)},
                "search": {"path": path, "overfetch": int(st.last_overfetch), "retried_queries": int(st.retried_queries),
                           "hint_retries": int(st.hint_retries), "max_abs_tc_err": float(st.max_abs_err),
                           "tail_ms": st.tail_ms_total / max(1, st.tail_ms_samples)}}
        if ms_e2e_pipe is not None:
            line["e2e"]["pipelined_value"] = w["b"] / (ms_e2e_pipe / steps * 1e-3)
        if not args.no_cpu_baseline and ctx.world == 1:
            t_step, xb, xq, cores = cpu_dense_sample(w, min(args.cpu_rows, w["n"]), 2, 1, None)
            rows = min(args.cpu_rows, w["n"])
            scale = w["n"] / rows
            cpu = {"value": w["b"] / (t_step * scale), "unit": "queries/s", "cores": cores, "kind": "port",
                   "sample": f"{rows} of {w['n']} rows x {w['b']} queries per step, time scaled x{scale:.1f}; torch {torch.__version__} sgemm, {cores} threads"}
            line["cpu_baseline"] = cpu
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), file=result_out, flush=True)


# ------------------------------------------------------------------------------------------------ latency workload (C5)
def run_latency(args, name, w, ctx, config, steps, warmup, result_out):
    """BASELINE configs[4]: single-query top-10 over a 50M x 768 corpus sharded over the GPUs, p50/p99 of the host-observed
    call latency (search + exchange + results visible), 1 000 timed calls after 100 warm-ups unless --steps says otherwise."""
    import numpy as np
    import torch
    from veritasfi_b200 import _native as N, synth
    from veritasfi_b200.sharded import make_sharded_dense

    if w["n"] / ctx.world * w["d"] * 2 > 150e9:
        raise SystemExit(f"{name} needs more GPUs: {w['n'] / ctx.world * w['d'] * 2 / 1e9:.0f} GB per shard")
    steps = steps if args.steps_given else 1000
    warmup = warmup if args.warmup_given else 100
    index, lo, hi = build_dense_index(ctx, w["n"], w["d"], SEED)
    n_local = hi - lo
    index.set_option(N.OPT_PROFILE, 1)
    searcher = make_sharded_dense(index, exchange=args.exchange, max_nq=w["b"], max_k=w["k"])
    q_dev = synth.dense_queries_torch(w["b"], w["d"], SEED, ctx.dev)
    parity = {"small": parity_small_dense(ctx, args.exchange)}
    lat = []

    def calls(k_steps):
        for _ in range(k_steps):
            t0 = time.perf_counter()
            # one batch in flight: the local search with the exchange enqueued right behind it (the rescoring kernel sends
            # its row to the peers itself), then one wait for both — no host round trip between search and exchange
            searcher.search_finish(searcher.search_begin(q_dev, w["k"]))
            torch.cuda.synchronize()
            lat.append((time.perf_counter() - t0) * 1e3)

    calls(warmup)
    lat.clear()
    index.stats(reset=True)
    sampler = ClockSampler(ctx.local_rank)
    if ctx.rank == 0:
        sampler.start()
        sampler.wait_first()
    launches0 = N.launch_count()
    ms_total = timed(ctx, calls, steps, host_bound=True)
    launches = N.launch_count() - launches0
    st = index.stats()
    clocks = sampler.stop() if ctx.rank == 0 else None
    ids, scores = searcher.search(q_dev, w["k"])
    torch.cuda.synchronize()
    parity["timed"] = parity_timed_rows(ctx, index, lo, hi, q_dev, ids, scores, n_check=1)
    kernel_ms = max_over_ranks(ctx, st.fused_ms_total / max(1, st.fused_ms_samples))
    p50 = max_over_ranks(ctx, float(np.percentile(lat, 50)))
    p99 = max_over_ranks(ctx, float(np.percentile(lat, 99)))
    if ctx.rank == 0:
        peaks = read_peaks()
        bytes_per_launch = float(n_local) * w["d"] * 2
        gbs = bytes_per_launch / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
        ms_step = ms_total / steps
        line = {"metric": "queries/sec", "value": w["b"] / (ms_step * 1e-3), "unit": "queries/s", "n_gpus": ctx.world, "steps": steps,
                "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic", "config": config,
                "latency_ms": {"p50": p50, "p99": p99, "mean": float(np.mean(lat)), "calls": steps, "warmups": warmup,
                               "what": "host-observed latency of one single-query call (search, exchange, results visible), max over ranks"},
                "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"], "traffic": None,
                             "kernel": "gemv_topk_kernel", "kernel_ms": kernel_ms,
                             "peak_source": f"{peaks['source']} MEASURED_PEAKS.json hbm_gbs (a pure read stream can exceed the copy figure)",
                             "algorithmic": f"(N/G)*d*2 = {bytes_per_launch:.4g} B per launch", "floor_ms": bytes_per_launch / (peaks['hbm'] * 1e9) * 1e3},
                "clocks": clocks,
                "e2e": {"value": w["b"] / (ms_step * 1e-3), "unit": "queries/s", "ms_per_step": ms_step, "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0, "how": "latency mode keeps the query and the result on the device; the timed call ends in a host synchronisation"},
                "gpu_launches": int(launches), "parity": parity, "cpu_baseline": None,
                "run": {"sharding": f"row-sharded over {ctx.world} GPU(s)" + (f", exchange={args.exchange}" if ctx.world > 1 else "")}}
        print(json.dumps(line), file=result_out, flush=True)


# ------------------------------------------------------------------------------------------------ hybrid workload (C4)
def build_hybrid(ctx, w, seed):
    """This rank's shards of the hybrid world: chunk rows, title rows, the doc-range postings (impacts from the GLOBAL df and
    avgdl, all-reduced once at setup), the global title -> chunk map and a fixed batch of token queries."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from veritasfi_b200 import synth
    from veritasfi_b200.bm25_compat import GpuPostings
    from veritasfi_b200.sharded import shard_bounds

    chunks, lo, hi = build_dense_index(ctx, w["n"], w["d"], seed)
    titles, tlo, thi = build_dense_index(ctx, w["n_ts"], w["d"], seed + 77)
    g = torch.Generator(device=ctx.dev)
    g.manual_seed(seed + 5)
    t2c = torch.randint(0, w["n"], (w["n_ts"],), generator=g, device=ctx.dev, dtype=torch.int64)   # same on every rank
    tok, doc, tf, dl = synth.zipf_postings_torch(hi - lo, w["vocab"], seed + 31 * ctx.rank, ctx.dev, mean_len=w["mean_len"])
    df = torch.bincount(tok, minlength=w["vocab"])
    tot = torch.tensor([float(dl.sum())], dtype=torch.float64, device=ctx.dev)
    if ctx.world > 1:
        dist.all_reduce(df)
        dist.all_reduce(tot)
    avgdl = float(tot.item()) / w["n"]
    indptr, indices, data = synth.bm25_impacts_torch(tok, doc, tf, dl, w["vocab"], w["n"], df, avgdl)
    del tok, doc, tf
    nnz = int(indices.numel())
    postings = GpuPostings.from_device(indptr, indices, data, hi - lo, id_offset=lo)
    df_h = df.cpu().numpy()
    del indptr, indices, data, df
    torch.cuda.empty_cache()
    queries = synth.bm25_queries(w["b"], w["vocab"], seed)
    toks, qptr = GpuPostings.pack_tokens(queries)
    return chunks, titles, t2c, postings, (toks, qptr), dict(lo=lo, hi=hi, nnz=nnz, df=df_h, queries=queries)


def parity_small_hybrid(ctx, exchange: str):
    """A small hybrid world sharded like the timed one (chunk rows, title rows, BM25 doc ranges), through the same retriever and
    exchange mode, against the stage-wise CPU oracle."""
    import numpy as np
    import torch
    from oracle import bm25 as obm, flat_ip, fusion as ofu
    from veritasfi_b200 import synth
    from veritasfi_b200.bm25_compat import GpuPostings, build_csc
    from veritasfi_b200.dense import DenseIndex
    from veritasfi_b200.multipath import MultiPathRetriever
    from veritasfi_b200.sharded import make_row_exchange, shard_bounds
    n, n_ts, d, B, L, k, V = 24_000, 5_000, 128, 24, 60, 20, 1500
    xb = synth.dense_corpus_np(n, d, 31)
    xq = synth.dense_queries_np(B, d, 31, xb)
    xt = synth.dense_corpus_np(n_ts, d, 32)
    t2c = np.random.default_rng(3).integers(0, n, size=n_ts).astype(np.int64)
    doc_ptr, toks = synth.zipf_postings(n, V, 4, mean_len=20)
    csc = build_csc(doc_ptr, toks, V)
    qs = synth.bm25_queries(B, V, 4)
    qs[0] = [V - 1]                               # a rare token: fewer than L matches -> zero-score filler must not be fused
    lo, hi = shard_bounds(n, ctx.world, ctx.rank)
    tlo, thi = shard_bounds(n_ts, ctx.world, ctx.rank)
    chunks = DenseIndex(d, device=ctx.dev, id_offset=lo)
    chunks.add(xb[lo:hi])
    titles = DenseIndex(d, device=ctx.dev, id_offset=tlo)
    titles.add(xt[tlo:thi])
    indptr, indices, data = csc
    # doc-range shard of the postings: per token, the postings with lo <= doc < hi, doc ids made local
    keep = (indices >= lo) & (indices < hi)
    tok_of = np.repeat(np.arange(V), np.diff(indptr))
    s_indptr = np.zeros(V + 1, dtype=np.int64)
    np.cumsum(np.bincount(tok_of[keep], minlength=V), out=s_indptr[1:])
    post = GpuPostings(s_indptr, (indices[keep] - lo).astype(np.int32), data[keep], hi - lo, id_offset=lo, device=ctx.dev.index)
    ex = make_row_exchange(ctx.dev, exchange, max_rows=3 * B, max_k=L)
    mp = MultiPathRetriever(chunks, titles, torch.from_numpy(t2c).to(ctx.dev), post, depth=L, sharded=ex)
    fi, fs, _ = mp.multipath_batch(torch.from_numpy(xq).to(ctx.dev), None, qs, k)
    torch.cuda.synchronize()
    D0, I0 = flat_ip.search(xq, xb, L)
    Dt, It = flat_ip.search(xq, xt, L)
    Ib, Sb = obm.retrieve(*csc, qs, n, L)
    oi, os_ = ofu.hybrid(np.stack([I0, It, Ib], axis=1), np.stack([D0, Dt, Sb], axis=1), t2c, 1, 2, 60.0, k)
    ok = bool((fi.cpu().numpy() == oi).all() and (fs.cpu().numpy() == os_).all())
    ok = all_ranks_ok(ctx, ok)
    if ex.exchange is not None:
        ex.exchange.close()
    for x in (chunks, titles, post):
        x.close()
    return {"ok": ok, "what": f"hybrid world {n} chunks / {n_ts} titles / V={V}, {B} queries, depth {L}, top-{k}, sharded over {ctx.world} GPU(s), "
                              f"exchange={exchange if ctx.world > 1 else 'none'}: fused ids and scores equal to the stage-wise oracle on every rank"}


def run_hybrid(args, name, w, ctx, config, steps, warmup, result_out):
    import numpy as np
    import torch
    from veritasfi_b200 import _native as N, synth
    from veritasfi_b200.multipath import MultiPathRetriever, fuse_hybrid
    from veritasfi_b200.sharded import make_row_exchange

    parity = {"small": parity_small_hybrid(ctx, args.exchange)}
    chunks, titles, t2c, postings, tokens, info = build_hybrid(ctx, w, SEED)
    for idx in (chunks, titles):
        idx.set_option(N.OPT_PROFILE, 1)
        idx.set_option(N.OPT_TAU_HINT, args.hint)
    postings.set_profile(True)
    B, L, k, P = w["b"], w["depth"], w["k"], 3
    ex = make_row_exchange(ctx.dev, args.exchange, max_rows=P * B, max_k=L)
    mp = MultiPathRetriever(chunks, titles, t2c, postings, depth=L, sharded=ex)
    q_dev = synth.dense_queries_torch(B, w["d"], SEED, ctx.dev)
    q_pin = torch.empty((B, w["d"]), dtype=torch.float32).pin_memory()
    q_pin.copy_(q_dev.cpu())
    q_stage = torch.empty_like(q_dev)
    out_i_pin = torch.empty((B, k), dtype=torch.int64).pin_memory()
    out_s_pin = torch.empty((B, k), dtype=torch.float32).pin_memory()

    def run_steps(k_steps):
        for _ in range(k_steps):
            mp.multipath_batch(q_dev, None, tokens, k)

    def run_e2e(k_steps):
        for _ in range(k_steps):
            q_stage.copy_(q_pin, non_blocking=True)
            fi, fs, _ = mp.multipath_batch(q_stage, None, tokens, k)
            out_i_pin.copy_(fi, non_blocking=True)
            out_s_pin.copy_(fs, non_blocking=True)
            torch.cuda.synchronize()

    run_steps(warmup)
    run_e2e(min(warmup, 3))
    chunks.stats(reset=True)
    titles.stats(reset=True)
    postings.stats(reset=True)
    sampler = ClockSampler(ctx.local_rank)
    if ctx.rank == 0:
        sampler.start()
        sampler.wait_first()
    launches0 = N.launch_count()
    ms_total = timed(ctx, run_steps, steps)
    launches = N.launch_count() - launches0
    clocks = sampler.stop() if ctx.rank == 0 else None
    sc, stt, sb = chunks.stats(), titles.stats(), postings.stats()
    ms_e2e = timed(ctx, run_e2e, steps, host_bound=True)
    # the fusion kernel alone, on lists of the timed shapes
    ids, scores = mp.path_lists(q_dev, None, tokens)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fuse_hybrid(ids, scores, t2c, 1, 2, k)
    e1.record()
    torch.cuda.synchronize()
    fuse_ms = e0.elapsed_time(e1) / 10
    k_chunks = max_over_ranks(ctx, sc.fused_ms_total / max(1, sc.fused_ms_samples))
    k_titles = max_over_ranks(ctx, stt.fused_ms_total / max(1, stt.fused_ms_samples))
    k_bm25 = max_over_ranks(ctx, sb.score_ms_total / max(1, sb.score_ms_samples))
    if ctx.rank == 0:
        peaks = read_peaks()
        ms_step = ms_total / steps
        n_loc, nt_loc = info["hi"] - info["lo"], titles.ntotal

        def dense_roof(kname, rows, ms):
            fl = 2.0 * B * rows * w["d"]
            tf = fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
            return {"kernel": kname, "bound": "tensor", "achieved": tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                    "frac": tf / peaks["tf_sustained"], "kernel_ms": ms, "algorithmic": f"2*B*rows*d = {fl:.4g} FLOP per launch ({rows} rows)"}
        pb = float(sb.postings_bytes)
        gbs = pb / (k_bm25 * 1e-3) / 1e9 if k_bm25 > 0 else 0.0
        fb = float(B) * P * L * 12 + B * k * 12
        rooflines = [
            dense_roof("dense_fused_pair_kernel<MODE_TOPK> (chunk path)", n_loc, k_chunks),
            dense_roof("dense_fused_pair_kernel<MODE_TOPK> (title path)", nt_loc, k_titles),
            {"kernel": "bm25_kernel", "bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
             "kernel_ms": k_bm25, "algorithmic": f"sum over queries and tokens of df*8 = {pb:.4g} B per launch (this shard's postings)"},
            {"kernel": "hybrid_fuse_kernel", "bound": "hbm", "achieved": fb / (fuse_ms * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
             "frac": fb / (fuse_ms * 1e-3) / 1e9 / peaks["hbm"], "kernel_ms": fuse_ms,
             "algorithmic": f"B*P*L*12 + B*k*12 = {fb:.4g} B per launch (latency-bound: two 1024-key block sorts per query)"},
        ]
        dominant = max(rooflines[:3], key=lambda r: r["kernel_ms"])
        roofline = dict(dominant)
        roofline["traffic"] = None
        roofline["peak_source"] = f"{peaks['source']} MEASURED_PEAKS.json"
        line = {"metric": "queries/sec", "value": B / (ms_step * 1e-3), "unit": "queries/s", "n_gpus": ctx.world, "steps": steps,
                "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "bf16 (dense paths), f32 (BM25 impacts, fused scores)", "data": "synthetic", "config": config,
                "roofline": roofline, "rooflines": rooflines, "clocks": clocks,
                "e2e": {"value": B / (ms_e2e / steps * 1e-3), "unit": "queries/s", "ms_per_step": ms_e2e / steps,
                        "h2d_bytes_per_step": B * w["d"] * 4 + int(tokens[0].nbytes + tokens[1].nbytes), "d2h_bytes_per_step": B * k * 12,
                        "how": "MultiPathRetriever.multipath_batch with a pinned H2D copy of the query embeddings, host token ids in, D2H of the fused ids/scores"},
                "gpu_launches": int(launches), "parity": parity, "cpu_baseline": None,
                "run": {"sharding": f"chunk rows, title rows and BM25 doc ranges sharded over {ctx.world} GPU(s)" +
                                    (f", one packed exchange of the three [B,{L}] lists per batch, exchange={args.exchange}" if ctx.world > 1 else ""),
                        "postings_nnz_local": info["nnz"], "tokens_per_batch": int(len(tokens[0]))}}
        print(json.dumps(line), file=result_out, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)   # default 100: ~1.7 s at C3, long enough for the power-cap governor to settle
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=list(WORKLOADS))
    ap.add_argument("--cpu-rows", type=int, default=250_000, help="rows of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full-cpu-step", action="store_true", help="reference arm: skip the one full-corpus step")
    ap.add_argument("--hint", type=int, default=1)
    ap.add_argument("--pair", type=int, default=None, help="VFI_OPT_CTA_PAIR (1 = single-CTA kernel, for comparison)")
    ap.add_argument("--tau-m", type=int, default=0, help="VFI_OPT_TAU_M (0 auto, 8/16/32: admission hint = m-th best of a row sample)")
    ap.add_argument("--no-small", action="store_true", help="batches of 9..64 queries on the pair kernel instead of the swapped-operand kernel")
    ap.add_argument("--e2e-stream", default="own", choices=["own", "caller"],
                    help="N=1 e2e host call: the library's pooled stream (default) or a stream supplied by the caller")
    ap.add_argument("--sync", action="store_true", help="one synchronous search (+ exchange) per step instead of two batches in flight")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: fused peer-memory push+merge kernel over NVLink, or NCCL all-gather + merge kernel")
    args = ap.parse_args()
    args.steps_given, args.warmup_given = args.steps is not None, args.warmup is not None
    w = dict(WORKLOADS[args.workload])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    steps = max(1, args.steps if args.steps is not None else 100)
    warmup = args.warmup if args.warmup is not None else 10
    warmup = max(3, warmup) if args.impl == "b200" else max(0, warmup)
    config = base_config(args.workload, w)

    if args.impl == "reference":
        if rank != 0:
            return
        reference_arm(args, args.workload, w, config, steps, warmup)
        return

    # stdout carries exactly ONE line, the JSON result: anything a library writes to file descriptor 1 meanwhile (NCCL prints
    # its version banner there) is sent to stderr instead
    sys.stdout.flush()
    result_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    ctx = Ctx()
    ctx.rank, ctx.world, ctx.local_rank = rank, world, local_rank
    ctx.dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=ctx.dev)
    runner = {"dense": run_dense, "latency": run_latency, "hybrid": run_hybrid}[w["kind"]]
    runner(args, args.workload, w, ctx, config, steps, warmup, result_out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
