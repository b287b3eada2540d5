#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "small_batch or exact or concurrent" 2>&1 | tail -5
show() {
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$1.json"))
    print("$1", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"].get("pipelined_value"), "e2e_ms", round(d["e2e"]["ms_per_step"],3), "e2e_kms", round(d["e2e"].get("kernel_ms",0),3), "ms", round(d["ms_per_step"],4), "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"],3), "kms", round(d["roofline"]["kernel_ms"],4), "parity", {k:v["ok"] for k,v in d["parity"].items()}, (d.get("search") or {}).get("tail_ms"))
except Exception as e: print("$1 failed", e)
PY
}
for wl in b16 b64; do
python bench.py --workload $wl --no-cpu-baseline 2>gpurun_out/e.err > gpurun_out/r2_bench_${wl}_1gpu.json || tail -5 gpurun_out/e.err; show r2_bench_${wl}_1gpu
done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/e.err > gpurun_out/x_c3_own.json || tail -5 gpurun_out/e.err; show x_c3_own
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-stream caller 2>gpurun_out/e.err > gpurun_out/x_c3_caller.json || tail -5 gpurun_out/e.err; show x_c3_caller
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/e.err > gpurun_out/x_c3_own2.json || tail -5 gpurun_out/e.err; show x_c3_own2
