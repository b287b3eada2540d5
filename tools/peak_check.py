#!/usr/bin/env python
"""Same-box calibration of the tensor roofline: cuBLAS bf16 GEMM sustained under the power cap next to K1 on C3,
each for a few seconds, with the SM clock sampled during the loop.  Box-to-box the power-capped clock differs by
20 %+, so a fraction of MEASURED_PEAKS.json (another box) is only meaningful beside this number.

    python tools/peak_check.py [seconds] [rows] [batch]
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import ClockSampler  # noqa: E402
from veritasfi_b200 import _native as N, synth  # noqa: E402
from veritasfi_b200.dense import DenseIndex  # noqa: E402

DEV = torch.device("cuda", 0)


def loop(fn, seconds):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s = ClockSampler(0)
    s.start()
    t_end = time.perf_counter() + seconds
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    e0.record()
    while time.perf_counter() < t_end:
        for _ in range(4):
            fn()
        n += 4
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, s.stop()


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    out = {}
    a = torch.randn(8192, 8192, device=DEV, dtype=torch.bfloat16)
    b = torch.randn(8192, 8192, device=DEV, dtype=torch.bfloat16)
    ms, clk = loop(lambda: torch.matmul(a, b), seconds)
    out["cublas_8192"] = {"ms": ms, "tflops": 2 * 8192 ** 3 / (ms * 1e-3) / 1e12, "clocks": clk}
    # the same shape as K1: [batch x 1024] x [1024 x rows-slice] (cuBLAS materialises the scores; 1M rows per call)
    d = 1024
    xs = torch.randn(1_000_000, d, device=DEV, dtype=torch.bfloat16)
    q = torch.randn(batch, d, device=DEV, dtype=torch.bfloat16)
    o = torch.empty(batch, 1_000_000, device=DEV, dtype=torch.bfloat16)
    ms, clk = loop(lambda: torch.matmul(q, xs.t(), out=o), seconds)
    out["cublas_qxT_1M"] = {"ms": ms, "tflops": 2 * batch * 1_000_000 * d / (ms * 1e-3) / 1e12, "clocks": clk}
    del a, b, xs, o
    idx = DenseIndex(d, store="bf16", device=DEV)
    idx.reserve(rows)
    for r0 in range(0, rows, 1 << 20):
        idx.add(synth.dense_corpus_torch(min(1 << 20, rows - r0), d, 7 + r0 // (1 << 20), DEV))
    qf = synth.dense_queries_torch(batch, d, 7, DEV)
    variants = [("k1", {N.OPT_TAU_HINT: 1})]
    variants += [(f"k1_pair{v}", {N.OPT_TAU_HINT: 1, N.OPT_CTA_PAIR: v}) for v in (1, 2)]
    only = os.environ.get("VFI_PEAK_ONLY")
    for name, opts in variants:
        if only and name not in only.split(","):
            continue
        for k_, v_ in opts.items():
            idx.set_option(k_, v_)
        idx.set_option(N.OPT_PROFILE, 1)
        idx.stats(reset=True)
        ms, clk = loop(lambda: idx.search_batch(qf, 100), seconds)
        st = idx.stats()
        kms = st.fused_ms_total / max(1, st.fused_ms_samples)
        out[name] = {"step_ms": ms, "kernel_ms": kms, "tflops": 2 * batch * rows * d / (kms * 1e-3) / 1e12, "clocks": clk,
                     "retried": int(st.retried_queries), "hint_retries": int(st.hint_retries)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
