"""Times the reference's ONLINE retrieval calls on the GPU kernels next to the CPU port:

  dense   faiss.IndexFlatIP.search(x[nq, d], k = 2048) for nq = 1 and 4 query strings
          (/root/reference/src/utils/ensembleRetriever.py:64-66 -> faissRetriever.py:37), fp32 store (faiss semantics)
  sparse  bm25s retrieve(k = N) read as [:bm25_k] (ensembleRetriever.py:189-190): eager top-2048 + lazy tail, and the
          full ranking for comparison

at C1 size (~10.9k x 1024) and at 1M x 1024.  One JSON object on stdout.  Run on a GPU box:
    python tools/online_call.py > gpurun_out/online_call.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import bm25 as obm, flat_ip  # noqa: E402
from veritasfi_b200 import _native as N, bm25_compat, faiss_compat, synth  # noqa: E402


def med(f, reps=20, warm=3):
    for _ in range(warm):
        f()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        f()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


def dense(n, d, out):
    xb = synth.dense_corpus_np(n, d, 5, bf16=False)
    xq = synth.dense_queries_np(4, d, 5, xb, bf16=False)
    index = faiss_compat.IndexFlatIP(d)
    index.add(xb)
    index.set_option(N.OPT_PROFILE, 1)
    for nq in (1, 4):
        ms = med(lambda: index.search(xq[:nq], 2048))
        index.stats(reset=True)
        for _ in range(10):
            index.search(xq[:nq], 2048)
        st = index.stats()
        kms = st.fused_ms_total / max(1, st.fused_ms_samples)
        cpu = med(lambda: flat_ip.search_faiss_like(xq[:nq], xb, 2048), reps=3, warm=1)
        out[f"dense_n{n}_nq{nq}"] = {
            "call_ms": ms, "device_ms": kms, "path": int(st.last_path), "corpus_bytes": n * d * 4,
            "achieved_GBs": n * d * 4 / (kms * 1e-3) / 1e9 if kms > 0 else None, "cpu_port_ms": cpu,
            "what": "IndexFlatIP.search(k=2048), fp32 rows, host buffers; device_ms = exact streaming scorer + radix select"}
    D, I = index.search(xq, 2048)
    D0, I0 = flat_ip.search(xq, xb, 2048)
    out[f"dense_n{n}_equals_oracle"] = bool((D == D0).all() and (I == I0).all())


def sparse(n_docs, vocab, out):
    doc_ptr, toks = synth.zipf_postings(n_docs, vocab, 6, mean_len=64)
    csc = bm25_compat.build_csc(doc_ptr, toks, vocab)
    eng = bm25_compat.BM25()
    eng.scores = {"data": csc[2], "indices": csc[1], "indptr": csc[0], "num_docs": n_docs}
    eng.corpus = [{"id": i, "text": ""} for i in range(n_docs)]
    qs = synth.bm25_queries(4, vocab, 6)

    def lazy():
        docs, sc = eng.retrieve([qs[0]], k=n_docs, return_as="tuple")
        return [d["id"] for d in docs[0][:10]], sc[0][:10]

    def full():
        docs, sc = eng.retrieve([qs[0]], k=n_docs, return_as="tuple")
        return np.asarray(sc[0])

    def cpu():
        s = obm.scores_numpy(*csc, qs[0], n_docs)
        order = np.argsort(-s, kind="stable")
        return order[:10], s[order[:10]]
    out[f"bm25_n{n_docs}"] = {
        "lazy_k_eq_N_read_top10_ms": med(lazy, reps=10), "full_ranking_ms": med(full, reps=5, warm=1),
        "cpu_port_ms": med(cpu, reps=3, warm=1), "postings": int(len(csc[1])),
        "what": "bm25s retrieve(k = N) then [:10]: eager top-2048 kernel path vs every doc ranked (radix sort + 12N bytes D2H) vs numpy add.at + argsort"}
    ids, sc = lazy()
    want = cpu()
    out[f"bm25_n{n_docs}_top10_equals_cpu"] = bool(list(want[0]) == list(ids))


def main():
    out = {}
    dense(10_907, 1024, out)
    sparse(10_907, 4096, out)
    dense(1_000_000, 1024, out)
    sparse(1_000_000, 65_536, out)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
