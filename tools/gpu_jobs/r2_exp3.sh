#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
show() {
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$1.json"))
    print("$1", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"].get("pipelined_value"), "ms", round(d["ms_per_step"],4), "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"],3), "kms", round(d["roofline"]["kernel_ms"],4), "parity", {k:v["ok"] for k,v in d["parity"].items()}, (d.get("search") or {}).get("tail_ms"))
    for r in d.get("rooflines", []): print("   ", r["kernel"], round(r["kernel_ms"],4), round(r["frac"],3))
except Exception as e: print("$1 failed", e)
PY
}
python bench.py --steps 20 --warmup 5 2>gpurun_out/e.err > gpurun_out/r2_bench_c3_1gpu_20steps.json || tail -5 gpurun_out/e.err; show r2_bench_c3_1gpu_20steps
python bench.py 2>gpurun_out/e.err > gpurun_out/r2_bench_c3_1gpu.json || tail -5 gpurun_out/e.err; show r2_bench_c3_1gpu
for wl in c2 c4s8; do
python bench.py --workload $wl --no-cpu-baseline 2>gpurun_out/e.err > gpurun_out/r2_bench_${wl}_1gpu.json || tail -5 gpurun_out/e.err; show r2_bench_${wl}_1gpu
done
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_arm.json 2>gpurun_out/e.err || tail -5 gpurun_out/e.err
cut -c1-1200 gpurun_out/r2_bench_reference_arm.json
python tools/online_call.py > gpurun_out/r2_online_call.json 2>gpurun_out/e.err; grep -E "device_ms|call_ms|lazy|full_rank|cpu_port" gpurun_out/r2_online_call.json | head -40
