#!/usr/bin/env python
"""bench.py — the retrieval hot path on B200, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c5|...] [--impl reference]

Metric (BASELINE.json): queries/sec at 10M x 1024, top-100, 1024-query batch (config C3); the corpus
is row-sharded over the N GPUs of the box (strong scaling: 10M rows in total for every N).
A step = one pass of the hot path over one query batch: prepare queries -> fused tcgen05 GEMM+top-k'
-> candidate reduction -> exact rescoring + certificate -> (N>1: NCCL all-gather + merge kernel).

  value     whole-job queries/s with queries and corpus resident in HBM
  e2e       the same through the reference-facing host call: pinned host queries in, host ids/scores
            out, copies inside the timed region
  roofline  the dominant kernel (dense_fused_kernel): algorithmic FLOPs 2*B*(N/G)*d per launch over its
            CUDA-event duration, against MEASURED_PEAKS.json (sustained bf16)
  cpu_baseline  the reference's CPU algorithm (oracle port of IndexFlatIP.search: blocked fp32 sgemm +
            running top-k) on the host cores, on a bounded row slice, scaled linearly

--impl reference runs only that CPU arm (rank 0) and prints the same line with "impl": "reference".
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
    # name: (total rows, dim, batch, k, store)
    "c3": dict(n=10_000_000, d=1024, b=1024, k=100, desc="10Mx1024 bf16 corpus, 1024-query batch, top-100 (BASELINE configs[2])"),
    "c2": dict(n=1_000_000, d=1024, b=256, k=100, desc="1Mx1024 bf16 corpus, 256-query batch, top-100 (BASELINE configs[1])"),
    "c5": dict(n=50_000_000, d=768, b=1, k=10, desc="50Mx768 bf16 corpus, single query, top-10 (BASELINE configs[4])"),
    "c3s8": dict(n=1_250_000, d=1024, b=1024, k=100, desc="one 1/8 row shard of C3 (1.25Mx1024 bf16), 1024-query batch, top-100"),
    "small": dict(n=200_000, d=1024, b=256, k=100, desc="200kx1024 smoke-size workload"),
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


def cpu_reference_arm(w, steps: int, warmup: int, rows_sample: int, threads: int | None):
    """The reference's CPU retrieval path (IndexFlatIP.search restated: blocked fp32 sgemm + running top-k)
    on the host cores, on a contiguous row slice of the same synthetic workload; q/s scaled to the full corpus."""
    import numpy as np
    import torch
    from oracle import flat_ip
    from veritasfi_b200 import synth

    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    rows = min(rows_sample, w["n"])
    xb = synth.dense_corpus_np(rows, w["d"], SEED, dup_frac=0.0, bf16=True)
    xq = synth.dense_queries_np(w["b"], w["d"], SEED, None, bf16=True)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        flat_ip.search_faiss_like(xq, xb, w["k"], threads=cores)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    t_step = sum(times) / len(times) * (w["n"] / rows)
    return dict(value=w["b"] / t_step, unit="queries/s", cores=cores, kind="port",
                sample=f"{rows} of {w['n']} rows x {w['b']} queries per step, time scaled x{w['n'] / rows:.1f}; "
                       f"torch {torch.__version__} sgemm, {cores} threads"), t_step


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)   # ~1.7 s at C3: long enough for the power-cap governor to settle
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=list(WORKLOADS))
    ap.add_argument("--cpu-rows", type=int, default=250_000, help="rows of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--hint", type=int, default=1)
    ap.add_argument("--sync", action="store_true",
                    help="N=1: one synchronous search per step instead of two batches in flight (search_begin/search_finish)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: fused peer-memory push+merge kernel over NVLink, or NCCL all-gather + merge kernel")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    steps, warmup = max(1, args.steps), max(3, args.warmup) if args.impl == "b200" else max(0, args.warmup)
    config = {"workload": f"{args.workload}: {w['desc']}", "corpus_rows": w["n"], "dim": w["d"], "batch": w["b"], "k": w["k"],
              "sharding": f"row-sharded over {world} GPU(s)" + (f", exchange={args.exchange}" if world > 1 else ""), "l2": "inputs exceed L2 (corpus shard >> 126 MB); no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return
        # exactly K timed steps after W warm-ups; a step is one pass over a bounded row slice (--cpu-rows), scaled to the corpus.
        # Bounded so the whole run ends within minutes whatever K is: at most ~120 s of CPU work.
        rows = args.cpu_rows
        est = 0.3 * (rows / 250_000) * (w["b"] / 1024)              # seconds per step on 16 host threads, measured on the pool
        while rows > 20_000 and est * (steps + warmup) > 120:
            rows //= 2
            est /= 2
        cpu, t_step = cpu_reference_arm(w, steps, warmup, rows, None)
        line = {"impl": "reference", "metric": "queries/sec", "value": cpu["value"], "unit": "queries/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": cpu,
                "e2e": {"value": cpu["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # stdout carries exactly ONE line, the JSON result: anything a library writes to file descriptor 1 meanwhile (NCCL prints
    # its version banner there) is sent to stderr instead
    sys.stdout.flush()
    result_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import numpy as np
    import torch
    import torch.distributed as dist
    from veritasfi_b200 import _native as N, synth
    from veritasfi_b200.dense import DenseIndex
    from veritasfi_b200.sharded import make_sharded_dense, shard_bounds

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    lo, hi = shard_bounds(w["n"], world, rank)
    n_local = hi - lo
    index = DenseIndex(w["d"], store="bf16", device=dev, id_offset=lo)
    index.reserve(n_local)
    chunk = 1 << 20
    for r0 in range(0, n_local, chunk):   # corpus generated per shard, on the device, seed + global chunk id
        r1 = min(n_local, r0 + chunk)
        index.add(synth.dense_corpus_torch(r1 - r0, w["d"], SEED + 1000 * rank + (r0 // chunk), dev))
    torch.cuda.synchronize()
    index.set_option(N.OPT_TAU_HINT, args.hint)
    index.set_option(N.OPT_PROFILE, 1)
    searcher = make_sharded_dense(index, exchange=args.exchange, max_nq=w["b"], max_k=w["k"])
    q_dev = synth.dense_queries_torch(w["b"], w["d"], SEED, dev)
    q_pin = torch.empty((w["b"], w["d"]), dtype=torch.float32).pin_memory()
    q_pin.copy_(q_dev.cpu())
    out_i_pin = torch.empty((w["b"], w["k"]), dtype=torch.int64).pin_memory()
    out_s_pin = torch.empty((w["b"], w["k"]), dtype=torch.float32).pin_memory()

    def step_device():
        return searcher.search(q_dev, w["k"])

    pipelined = world == 1 and not args.sync
    config["pipeline"] = ("two batches in flight: vfi_index_search_begin(i+1) is enqueued before vfi_index_search_finish(i) reads the "
                          "certificate flag of batch i" if pipelined else "one synchronous search (+ exchange) per step")

    def run_device_steps(k_steps):
        """K steps of the hot path with inputs resident in HBM.  N=1: a serving loop keeps two batches in flight so the GPU does
        not idle while the host looks at the certificate flag; every batch is still certified (and repaired if needed) inside
        the timed region.  N>1: local search, then the exchange, synchronously per batch."""
        if not pipelined:
            for _ in range(k_steps):
                step_device()
            return
        prev = None
        for _ in range(k_steps):
            t = index.search_begin(q_dev, w["k"])
            if prev is not None:
                index.search_finish(prev)
            prev = t
        index.search_finish(prev)

    def step_e2e():
        if world == 1:   # the reference-facing host call of the C ABI: H2D, search, D2H inside
            index.search_host_into(q_pin.data_ptr(), w["b"], w["k"], out_s_pin.data_ptr(), out_i_pin.data_ptr())
        else:
            qd = q_pin.to(dev, non_blocking=True)
            ids, scores = searcher.search(qd, w["k"])
            out_i_pin.copy_(ids, non_blocking=True)
            out_s_pin.copy_(scores, non_blocking=True)
            torch.cuda.synchronize()

    def timed(fn, k_steps, block=False):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        if block:
            fn(k_steps)
        else:
            for _ in range(k_steps):
                fn()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), 0.0)
        # host-synchronous steps (e2e) are bounded below by wall time; use the larger of the two clocks
        ms = max(ms, wall * 1e3) if fn is step_e2e else ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    run_device_steps(warmup)
    for _ in range(min(warmup, 3)):
        step_e2e()
    if os.environ.get("VFI_BENCH_BREAKDOWN"):
        from veritasfi_b200 import sharded as sh
        from veritasfi_b200.dense import merge_topk

        def tick(label, fn, acc):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn()
            torch.cuda.synchronize()
            acc[label] = acc.get(label, 0.0) + (time.perf_counter() - t0) * 1e3
            return r
        acc = {}
        for _ in range(5):
            ids_l, sc_l = tick("local_search", lambda: index.search_batch(q_dev, w["k"]), acc)
            if world > 1:
                mine = tick("pack", lambda: sh.pack(sc_l, ids_l), acc)
                flat = torch.empty((world * mine.shape[0], mine.shape[1]), dtype=mine.dtype, device=dev)
                tick("all_gather", lambda: dist.all_gather_into_tensor(flat, mine), acc)
                gs, gi = tick("unpack", lambda: sh.unpack(flat.view(world, mine.shape[0], mine.shape[1]), w["k"]), acc)
                tick("merge", lambda: merge_topk(gs, gi, w["k"]), acc)
            tick("h2d_queries", lambda: q_pin.to(dev, non_blocking=True), acc)
            tick("d2h_results", lambda: (out_i_pin.copy_(ids_l, non_blocking=True), out_s_pin.copy_(sc_l, non_blocking=True)), acc)
        print(f"[breakdown rank {rank}] " + "  ".join(f"{k}={v / 5:.3f}ms" for k, v in acc.items()), file=sys.stderr, flush=True)
    index.stats(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    launches0 = N.launch_count()
    ms_total = timed(run_device_steps, steps, block=True)
    launches = N.launch_count() - launches0
    st = index.stats()
    clocks = sampler.stop() if rank == 0 else None
    index.stats(reset=True)
    ms_e2e = timed(step_e2e, steps)
    st_e2e = index.stats()
    ids, scores = step_device()
    torch.cuda.synchronize()

    kernel_ms = st.fused_ms_total / max(1, st.fused_ms_samples)
    kt = torch.tensor([kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(kt, op=dist.ReduceOp.MAX)
    kernel_ms = float(kt.item())

    if rank == 0:
        peaks = read_peaks()
        ms_step = ms_total / steps
        flops_per_launch = 2.0 * w["b"] * n_local * w["d"]
        achieved_tf = flops_per_launch / (kernel_ms * 1e-3) / 1e12 if kernel_ms > 0 else 0.0
        tensor_bound = w["b"] >= 128
        if tensor_bound:
            roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                        "frac": achieved_tf / peaks["tf_sustained"],
                        "traffic": read_traffic(args.workload) if world == 1 else None,
                        "kernel": "dense_fused_pair_kernel<MODE_TOPK>", "kernel_ms": kernel_ms,
                        "peak_source": f"{peaks['source']} MEASURED_PEAKS.json bf16_tflops_sustained",
                        "algorithmic": f"2*B*(N/G)*d = {flops_per_launch:.4g} FLOP per launch; (N/G)*d*2 = {float(n_local) * w['d'] * 2:.4g} B",
                        "traffic_source": "ncu dram__bytes_read+write per launch, profiles/ncu_traffic.json"}
        else:
            bytes_per_launch = float(n_local) * w["d"] * 2
            gbs = bytes_per_launch / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
            roofline = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                        "traffic": None, "kernel": "gemv_topk_kernel", "kernel_ms": kernel_ms,
                        "peak_source": f"{peaks['source']} MEASURED_PEAKS.json hbm_gbs",
                        "algorithmic": f"(N/G)*d*2 = {bytes_per_launch:.4g} B per launch"}
        line = {"metric": "queries/sec", "value": w["b"] / (ms_step * 1e-3), "unit": "queries/s", "n_gpus": world, "steps": steps,
                "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic", "config": config, "roofline": roofline, "clocks": clocks,
                "e2e": {"value": w["b"] / (ms_e2e / steps * 1e-3), "unit": "queries/s", "ms_per_step": ms_e2e / steps,
                        "h2d_bytes_per_step": w["b"] * w["d"] * 4, "d2h_bytes_per_step": w["b"] * w["k"] * 12,
                        "kernel_ms": st_e2e.fused_ms_total / max(1, st_e2e.fused_ms_samples),
                        "tail_ms": st_e2e.tail_ms_total / max(1, st_e2e.tail_ms_samples)},
                "gpu_launches": int(launches),
                "search": {"path": int(st.last_path), "overfetch": int(st.last_overfetch), "retried_queries": int(st.retried_queries),
                           "hint_retries": int(st.hint_retries), "max_abs_tc_err": float(st.max_abs_err),
                           "tail_ms": st.tail_ms_total / max(1, st.tail_ms_samples)}}
        if not args.no_cpu_baseline and world == 1:
            cpu, _ = cpu_reference_arm(w, 2, 1, args.cpu_rows, None)
            line["cpu_baseline"] = cpu
        elif world > 1:
            line["cpu_baseline"] = None
        print(json.dumps(line), file=result_out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
