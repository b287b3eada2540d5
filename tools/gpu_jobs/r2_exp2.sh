#!/bin/bash
mkdir -p gpurun_out
./tools/ubench/cluster_occupancy | tee gpurun_out/r2_cluster_occupancy.txt
python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
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
for wl in c2 c3s8 c4s8; do
python bench.py --workload $wl --no-cpu-baseline 2>gpurun_out/e.err > gpurun_out/r2_bench_${wl}_1gpu.json || tail -5 gpurun_out/e.err; show r2_bench_${wl}_1gpu
done
python tools/prof_step.py --workload c4s8 --steps 1 --warmup 1 > gpurun_out/plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'bm25_kernel' -s 1 -c 1 -o gpurun_out/r2_bm25_c4s8 \
    python tools/prof_step.py --workload c4s8 --steps 1 --warmup 1 > gpurun_out/ncu_bm25.log 2>&1
ls -la gpurun_out/r2_bm25_c4s8.ncu-rep
