#!/usr/bin/env python
"""Measurements of the other BASELINE configs on one GPU (evidence for DESIGN.md / profiles/; the driver's
contract line is bench.py).  Each prints one JSON line.

    python tools/bench_extra.py c5s    # latency mode: one shard (1/8) of 50M x 768, single query, top-10, p50/p99
    python tools/bench_extra.py c4s    # hybrid 3-path: one shard (1/8) of 5M chunks: dense + title dense + BM25 + RRF
    python tools/bench_extra.py bm25   # BM25 kernel alone at a larger shard, GB/s of postings vs HBM peak
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import read_peaks  # noqa: E402
from veritasfi_b200 import _native as N, fusion as F, synth  # noqa: E402
from veritasfi_b200.bm25_compat import GpuPostings, build_csc  # noqa: E402
from veritasfi_b200.dense import DenseIndex  # noqa: E402
from veritasfi_b200.multipath import MultiPathRetriever  # noqa: E402

DEV = torch.device("cuda", 0)


def build_dense(n, d, seed, store="bf16"):
    idx = DenseIndex(d, store=store, device=DEV)
    idx.reserve(n)
    chunk = 1 << 20
    for r0 in range(0, n, chunk):
        idx.add(synth.dense_corpus_torch(min(chunk, n - r0), d, seed + r0 // chunk, DEV))
    torch.cuda.synchronize()
    return idx


def c5s():
    peaks = read_peaks()
    n, d, k = 50_000_000 // 8, 768, 10
    idx = build_dense(n, d, 5000)
    idx.set_option(N.OPT_PROFILE, 1)
    qs = synth.dense_queries_torch(1100, d, 5000, DEV).cpu().numpy()
    for i in range(100):
        idx.search_host(qs[i:i + 1], k)
    idx.stats(reset=True)
    lat = []
    for i in range(100, 1100):
        t0 = time.perf_counter()
        idx.search_host(qs[i:i + 1], k)
        lat.append((time.perf_counter() - t0) * 1e3)
    st = idx.stats()
    lat.sort()
    kms = st.fused_ms_total / max(1, st.fused_ms_samples)
    gbs = n * d * 2 / (kms * 1e-3) / 1e9
    print(json.dumps({"workload": "c5s: one 1/8 shard (6.25M x 768 bf16) of BASELINE configs[4], single query, top-10, host call",
                      "metric": "latency_ms", "p50": lat[500], "p99": lat[990], "mean": sum(lat) / len(lat), "calls": 1000,
                      "path": int(st.last_path), "retried_queries": int(st.retried_queries),
                      "roofline": {"bound": "hbm", "kernel": "gemv_topk_kernel", "kernel_ms": kms, "achieved": gbs, "peak": peaks["hbm"],
                                   "unit": "GB/s", "frac": gbs / peaks["hbm"], "algorithmic": f"{n}*{d}*2 B per launch"}}))


def make_postings(n_docs, n_vocab, seed, mean_len=128):
    t0 = time.time()
    doc_ptr, toks = synth.zipf_postings(n_docs, n_vocab, seed, mean_len=mean_len)
    csc = build_csc(doc_ptr, toks, n_vocab)
    return csc, time.time() - t0


def bm25(n_docs=1_000_000, n_vocab=262_144, nq=1024, k=50):
    peaks = read_peaks()
    csc, t_build = make_postings(n_docs, n_vocab, 4000)
    gp = GpuPostings(*csc, n_docs)
    gp.set_profile(True)
    qs = synth.bm25_queries(nq, n_vocab, 4000)
    for _ in range(2):
        gp.search(qs, k)
    gp.stats(reset=True)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        ids, scores = gp.search(qs, k)
    wall = (time.perf_counter() - t0) / reps * 1e3
    st = gp.stats()
    kms = st.score_ms_total / max(1, st.score_ms_samples)
    gbs = st.postings_bytes / (kms * 1e-3) / 1e9
    print(json.dumps({"workload": f"bm25: {n_docs} docs, V={n_vocab}, nnz={len(csc[1])}, {nq} queries of 4-16 tokens, top-{k}",
                      "metric": "queries/sec", "value": nq / (wall * 1e-3), "ms_per_batch": wall, "index_build_s": t_build,
                      "roofline": {"bound": "hbm", "kernel": "bm25_kernel", "kernel_ms": kms, "achieved": gbs, "peak": peaks["hbm"],
                                   "unit": "GB/s", "frac": gbs / peaks["hbm"],
                                   "algorithmic": f"sum over query tokens df*8 = {st.postings_bytes} B per launch"}}))


def c4s():
    peaks = read_peaks()
    n, n_ts, d, B, L, k, V = 5_000_000 // 8, 1_000_000 // 8, 1024, 1024, 200, 50, 262_144
    chunks = build_dense(n, d, 4100)
    titles = build_dense(n_ts, d, 4200)
    for ix in (chunks, titles):
        ix.set_option(N.OPT_TAU_HINT, 1)
        ix.set_option(N.OPT_PROFILE, 1)
    t2c = torch.randint(0, n, (n_ts,), device=DEV, generator=torch.Generator(device=DEV).manual_seed(7))
    csc, t_build = make_postings(n, V, 4300)
    gp = GpuPostings(*csc, n)
    gp.set_profile(True)
    mp = MultiPathRetriever(chunks, titles, t2c, gp, depth=L)
    q = synth.dense_queries_torch(B, d, 4100, DEV)
    toks = GpuPostings.pack_tokens(synth.bm25_queries(B, V, 4300))
    for _ in range(2):
        mp.multipath_batch(q, None, toks, k, fusion="rrf")
    torch.cuda.synchronize()
    for o in (chunks, titles):
        o.stats(reset=True)
    gp.stats(reset=True)
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        ids, scores, _ = mp.multipath_batch(q, None, toks, k, fusion="rrf")
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e3
    sc, stt, sb = chunks.stats(), titles.stats(), gp.stats()
    kc = sc.fused_ms_total / max(1, sc.fused_ms_samples)
    kt = stt.fused_ms_total / max(1, stt.fused_ms_samples)
    kb = sb.score_ms_total / max(1, sb.score_ms_samples)
    # fusion alone
    lists = torch.randint(0, n, (B, 3, L), device=DEV)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        F.rrf(lists, k)
    e1.record()
    torch.cuda.synchronize()
    rrf_ms = e0.elapsed_time(e1) / 20
    print(json.dumps({
        "workload": f"c4s: one 1/8 shard of BASELINE configs[3]: {n} chunks + {n_ts} title vectors x {d} bf16, BM25 V={V} nnz={len(csc[1])}, "
                    f"B={B}, depth {L}, RRF-60, top-{k}",
        "metric": "queries/sec", "value": B / (wall * 1e-3), "ms_per_batch": wall,
        "kernels": {
            "dense_chunks": {"ms": kc, "tflops": 2.0 * B * n * d / (kc * 1e-3) / 1e12, "frac_of_sustained": 2.0 * B * n * d / (kc * 1e-3) / 1e12 / peaks["tf_sustained"]},
            "dense_titles": {"ms": kt, "tflops": 2.0 * B * n_ts * d / (kt * 1e-3) / 1e12, "frac_of_sustained": 2.0 * B * n_ts * d / (kt * 1e-3) / 1e12 / peaks["tf_sustained"]},
            "bm25": {"ms": kb, "gbs": sb.postings_bytes / (kb * 1e-3) / 1e9, "frac_of_hbm": sb.postings_bytes / (kb * 1e-3) / 1e9 / peaks["hbm"], "postings_bytes": sb.postings_bytes},
            "rrf": {"ms": rrf_ms, "gbs": (B * 3 * L * 8 + B * k * 12) / (rrf_ms * 1e-3) / 1e9},
        }}))


if __name__ == "__main__":
    {"c5s": c5s, "c4s": c4s, "bm25": bm25}[sys.argv[1]]()
