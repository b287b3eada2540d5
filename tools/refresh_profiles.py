#!/usr/bin/env python
"""Rebuild the machine-made sections of profiles/r2_ncu_summary.md from what a `tools/gpu_jobs/r2_final_1gpu.sh` run left in
gpurun_out/: step shares from the ncu launch lists, the `--set full` summaries of K1 (C3, C2 + its tail) and of the small-batch
kernel, the per-kernel SASS counts of the built library; refreshes profiles/ncu_traffic.json and copies the launch lists.
    python tools/refresh_profiles.py"""
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402


def step_share(csv_path, last_steps=1):
    """Kernels of the LAST step in the launch list (the step = everything from the last prep_queries of the first search of a step)."""
    rows = []
    with open(csv_path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        if "vfi::" not in name:
            continue
        short = re.sub(r"\(.*$", "", name.replace("void ", ""))
        rows.append((short, float(r["Metric Value"]) / 1e3))
    return rows


def last_step(rows, first_kernel_prefix, n_first):
    """Slice the final step: it starts at the n_first-th occurrence (from the end) of a kernel whose name starts with the prefix."""
    idx = [i for i, (n, _) in enumerate(rows) if n.startswith(first_kernel_prefix)]
    start = idx[-n_first]
    return rows[start:]


def table(rows):
    tot = sum(t for _, t in rows)
    out = ["| kernel | device time | share |", "|---|---|---|"]
    for n, t in rows:
        out.append(f"| `{n}` | {t:.1f} us | {100 * t / tot:.1f} % |")
    out.append(f"| total | {tot:.1f} us | |")
    return "\n".join(out), tot


def ncu_bullets(rep, extra=()):
    old = sys.argv
    sys.argv = ["ncu_summary.py", rep, *extra]
    import io
    from contextlib import redirect_stdout
    buf = io.StringIO()
    with redirect_stdout(buf):
        ncu_summary.main()
    sys.argv = old
    return buf.getvalue().strip()


def raw_metric(rep, kernel_regex, metric):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    ki, mi = hdr.index("Kernel Name"), hdr.index(metric)
    for r in rows[2:]:
        if re.search(kernel_regex, r[ki]):
            return float(r[mi].replace(",", "")), rows[1][mi]
    return None, None


def sass_table():
    lib = os.path.join(ROOT, "veritasfi_b200", "_lib", "libvfi.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True, text=True).stdout.split("\n")
    blocks = re.split(r"\n\s*Function : \S+\n", "\n" + txt)[1:]
    cols = [("UTCHMMA", "UTCHMMA (tcgen05.mma)"), ("LDTM", "LDTM (tcgen05.ld)"), ("UTMALDG", "UTMALDG (TMA tensor load)"),
            ("UBLKCP", "UBLKCP (bulk copy)"), ("UTCBAR", "UTCBAR (tcgen05.commit)"), ("DFMA", "DFMA"), ("F2F.F64.F32", "F2F.F64.F32"),
            ("HMMA", "HMMA (legacy mma.sync)")]
    out = ["| kernel | " + " | ".join(c[1] for c in cols) + " |", "|" + "---|" * (len(cols) + 1)]
    keep = ("dense_fused", "dense_small", "exact_scores", "rescore_finalize", "bm25_kernel", "exchange_", "gemv_topk")
    rows = []
    for n, b in zip(names, blocks):
        short = re.sub(r"\(.*$", "", n.replace("void ", ""))
        if not any(k in short for k in keep):
            continue
        counts = []
        for key, _ in cols:
            if key == "HMMA":
                counts.append(len(re.findall(r"\bHMMA\b", b)))
            else:
                counts.append(len(re.findall(r"\b" + re.escape(key), b)))
        rows.append((short, counts))
    for short, counts in sorted(rows):
        out.append(f"| `{short}` | " + " | ".join(str(c) for c in counts) + " |")
    return "\n".join(out)


def main():
    md_path = os.path.join(PROF, "r2_ncu_summary.md")
    md = open(md_path).read()
    bench = {w: json.load(open(os.path.join(OUT, f"r2_bench_{w}_1gpu.json"))) for w in ("c3", "c2", "c4s8")}

    secs = {}
    r = last_step(step_share(os.path.join(OUT, "r2_launches_c3.csv")), "vfi::prep_queries_kernel", 1)
    t3, tot3 = table(r)
    k1 = max(t for n, t in r if "dense_fused_pair_kernel<0>" in n)
    live = bench["c3"]
    secs["c3"] = (t3 + f"\n\nK1's share under ncu {100 * k1 / tot3:.1f} %; live (CUDA events inside `bench.py`, `profiles/r2_bench_c3_1gpu.json`): "
                  f"{live['roofline']['kernel_ms']:.2f} / {live['ms_per_step']:.2f} ms = {100 * live['roofline']['kernel_ms'] / live['ms_per_step']:.1f} %.")
    r = last_step(step_share(os.path.join(OUT, "r2_launches_c2.csv")), "vfi::prep_queries_kernel", 1)
    t2, tot2 = table(r)
    live = bench["c2"]
    secs["c2"] = (t2 + f"\n\nLive: K1 {live['roofline']['kernel_ms']:.3f} ms of a {live['ms_per_step']:.3f} ms step "
                  f"({100 * live['roofline']['kernel_ms'] / live['ms_per_step']:.0f} %; the live step also carries launch gaps).")
    r = last_step(step_share(os.path.join(OUT, "r2_launches_c4s8.csv")), "vfi::prep_queries_kernel", 2)
    t4, _ = table(r)
    secs["c4s8"] = t4

    def replace_between(text, start_marker, end_marker, body):
        a = text.index(start_marker) + len(start_marker)
        b = text.index(end_marker, a)
        return text[:a] + "\n" + body + "\n\n" + text[b:]

    md = replace_between(md, "### C3 (10M x 1024 bf16, 1024 queries, top-100), `profiles/r2_launches_c3.csv`", "### C2 (1M x 1024, 256 queries)", secs["c3"])
    md = replace_between(md, "### C2 (1M x 1024, 256 queries), `profiles/r2_launches_c2.csv`", "### C4 one-eighth shard", secs["c2"])
    md = replace_between(md, "### C4 one-eighth shard (625k chunks + 125k titles + BM25, 1024 queries, depth 200), `profiles/r2_launches_c4s8.csv`",
                         "## K1 `dense_fused_pair_kernel<TOPK>` at C3", secs["c4s8"])

    rep3, rep2, reps = (os.path.join(OUT, n) for n in ("r2_k1_c3.ncu-rep", "r2_k1_c2.ncu-rep", "r2_k1s_b64.ncu-rep"))
    extra = ["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
             "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__waves_per_multiprocessor",
             "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
    b3 = ncu_bullets(rep3)
    rd3, _ = raw_metric(rep3, "dense_fused_pair_kernel<0>", "dram__bytes_read.sum")
    wr3, _ = raw_metric(rep3, "dense_fused_pair_kernel<0>", "dram__bytes_write.sum")
    md = replace_between(md, "## K1 `dense_fused_pair_kernel<TOPK>` at C3 (`gpurun_out/r2_k1_c3.ncu-rep`; this is `roofline.traffic` of the default bench line)",
                         "## K1 at C2, its sample pass, and the tail K1c + K2", b3 +
                         "\n\nEvery corpus tile leaves HBM once (DRAM traffic / 20.48 GB of corpus below); the L2->SM fill is 8 x the corpus: every (corpus tile, "
                         "query-tile-pair) item re-fetches its query block from L2.")
    b2 = ncu_bullets(rep2, extra)
    md = replace_between(md, "## K1 at C2, its sample pass, and the tail K1c + K2 (`gpurun_out/r2_k1_c2.ncu-rep`)",
                         "## K1s `dense_small_kernel` at 64 queries", b2)
    bs = ncu_bullets(reps)
    md = replace_between(md, "## K1s `dense_small_kernel` at 64 queries over 1M x 1024 (`gpurun_out/r2_k1s_b64.ncu-rep`) — new in round 2",
                         "## Exact streaming scorer", bs)
    a = md.index("## SASS per kernel")
    b = md.index("No `HMMA` (legacy `mma.sync`) anywhere")
    md = md[:a] + "## SASS per kernel (`cuobjdump -sass veritasfi_b200/_lib/libvfi.so`, instruction occurrences in the code, sm_100a)\n" + sass_table() + "\n\n" + md[b:]
    open(md_path, "w").write(md)

    # traffic json
    tj = json.load(open(os.path.join(PROF, "ncu_traffic.json")))

    def units(v, u):
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
    for key, rep, rx in (("c3", rep3, "dense_fused_pair_kernel<0>"), ("c2", rep2, "dense_fused_pair_kernel<0>"), ("b64", reps, "dense_small_kernel")):
        rd, ru = raw_metric(rep, rx, "dram__bytes_read.sum")
        wr, wu = raw_metric(rep, rx, "dram__bytes_write.sum")
        if rd is not None:
            tj[key]["dram_bytes_read"] = units(rd, ru)
            tj[key]["dram_bytes_write"] = units(wr, wu)
    json.dump(tj, open(os.path.join(PROF, "ncu_traffic.json"), "w"), indent=1)
    for w in ("c3", "c2", "c4s8"):
        shutil.copy(os.path.join(OUT, f"r2_launches_{w}.csv"), os.path.join(PROF, f"r2_launches_{w}.csv"))
    print("refreshed", md_path)


if __name__ == "__main__":
    main()
