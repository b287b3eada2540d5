#!/bin/bash
# round 2, one GPU: tests, smoke, the default bench, the other workloads, the online-call timing
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 2>gpurun_out/r2_c3.err > gpurun_out/r2_bench_c3_1gpu.json; tail -c 1500 gpurun_out/r2_bench_c3_1gpu.json; echo
for wl in c2 b16 b64 b128 c5s8 c4s8; do
  python bench.py --workload $wl --no-cpu-baseline 2>gpurun_out/r2_$wl.err > gpurun_out/r2_bench_${wl}_1gpu.json || tail -5 gpurun_out/r2_$wl.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench_${wl}_1gpu.json"))
    print("$wl", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"],3), "kms", round(d["roofline"]["kernel_ms"],4), "parity", {k:v["ok"] for k,v in d["parity"].items()}, d.get("latency_ms"))
    for r in d.get("rooflines", []): print("   ", r["kernel"], round(r["kernel_ms"],4), round(r["frac"],3))
except Exception as e: print("$wl failed", e)
PY
done
python tools/online_call.py > gpurun_out/r2_online_call.json 2>gpurun_out/r2_online.err; cat gpurun_out/r2_online_call.json | head -80
