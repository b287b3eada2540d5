#!/bin/bash
# experiments: packed BM25 postings; admission-hint density at C2 / C3 shard
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_reference_boundary.py -m gpu -x -q 2>&1 | tail -4
show() {
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$1.json"))
    print("$1", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"],3), "kms", round(d["roofline"]["kernel_ms"],4), "parity", {k:v["ok"] for k,v in d["parity"].items()}, (d.get("search") or {}).get("tail_ms"))
    for r in d.get("rooflines", []): print("   ", r["kernel"], round(r["kernel_ms"],4), round(r["frac"],3))
except Exception as e: print("$1 failed", e)
PY
}
python bench.py --workload c4s8 --no-cpu-baseline 2>gpurun_out/e.err > gpurun_out/r2_bench_c4s8_1gpu.json || tail -5 gpurun_out/e.err; show r2_bench_c4s8_1gpu
for m in 0 16 32; do
  python bench.py --workload c2 --no-cpu-baseline --tau-m $m 2>gpurun_out/e.err > gpurun_out/x_c2_m$m.json || tail -5 gpurun_out/e.err; show x_c2_m$m
done
for m in 0 16; do
  python bench.py --workload c3s8 --no-cpu-baseline --tau-m $m 2>gpurun_out/e.err > gpurun_out/x_c3s8_m$m.json || tail -5 gpurun_out/e.err; show x_c3s8_m$m
done
