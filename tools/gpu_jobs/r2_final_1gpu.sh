#!/bin/bash
# round 2, one GPU, at HEAD: the driver's sequence (tests, smoke, bench), every workload, then the ncu evidence for profiles/
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -12
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
show() {
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$1.json"))
    print("$1", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"].get("pipelined_value"), "ms", round(d["ms_per_step"],4), "roof", d["roofline"]["kernel"][:28], round(d["roofline"]["frac"],3), "kms", round(d["roofline"]["kernel_ms"],4), "parity", {k:v["ok"] for k,v in d["parity"].items()}, (d.get("search") or {}).get("tail_ms"), d.get("latency_ms"))
    for r in d.get("rooflines", []): print("   ", r["kernel"], round(r["kernel_ms"],4), round(r["frac"],3))
except Exception as e: print("$1 failed", e)
PY
}
python bench.py --steps 20 --warmup 5 2>gpurun_out/e.err > gpurun_out/r2_bench_c3_1gpu_20steps.json || tail -5 gpurun_out/e.err; show r2_bench_c3_1gpu_20steps
python bench.py 2>gpurun_out/e.err > gpurun_out/r2_bench_c3_1gpu.json || tail -5 gpurun_out/e.err; show r2_bench_c3_1gpu
for wl in c2 c3s8 c4s8 c5s8 b16 b64 b128; do
python bench.py --workload $wl --no-cpu-baseline 2>gpurun_out/e.err > gpurun_out/r2_bench_${wl}_1gpu.json || tail -5 gpurun_out/e.err; show r2_bench_${wl}_1gpu
done
for wl in b16 b64; do
python bench.py --workload $wl --no-cpu-baseline --no-small 2>gpurun_out/e.err > gpurun_out/r2_bench_${wl}_pairkernel_1gpu.json || tail -5 gpurun_out/e.err; show r2_bench_${wl}_pairkernel_1gpu
done
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_arm.json 2>gpurun_out/e.err || tail -5 gpurun_out/e.err
VFI_TRACE_HOST=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2_trace_host.err >/dev/null; grep "vfi trace" gpurun_out/r2_trace_host.err | tail -12
# per-operation device times of a batch (one event behind every stream operation)
: > gpurun_out/r2_step_trace.txt
for wl in c3s8 c2 c3; do
  VFI_TRACE_STEPS=1 python tools/trace_steps.py --workload $wl --profile 0 --steps 8 2>&1 | grep -A12 "^--- $wl" | tail -6 | sed "s/^/[$wl] /" >> gpurun_out/r2_step_trace.txt
done
cat gpurun_out/r2_step_trace.txt
# ---- ncu: launch lists (shares of a step) and full captures of the kernels profiles/ quotes
for wl in c3 c2 c4s8; do
  python tools/prof_step.py --workload $wl --steps 2 --warmup 2 > gpurun_out/plain_$wl.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_$wl.csv \
      python tools/prof_step.py --workload $wl --steps 2 --warmup 2 > gpurun_out/ncu_$wl.log 2>&1
done
python tools/prof_step.py --workload c3 --steps 1 --warmup 1 > gpurun_out/plain_c3b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'dense_fused_pair_kernel' -s 3 -c 1 -o gpurun_out/r2_k1_c3 \
    python tools/prof_step.py --workload c3 --steps 1 --warmup 1 > gpurun_out/ncu_k1c3.log 2>&1
python tools/prof_step.py --workload c2 --steps 1 --warmup 1 > gpurun_out/plain_c2b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'dense_fused_pair_kernel|rescore_finalize|cand_reduce' -s 3 -c 3 -o gpurun_out/r2_k1_c2 \
    python tools/prof_step.py --workload c2 --steps 1 --warmup 1 > gpurun_out/ncu_k1c2.log 2>&1
python tools/prof_step.py --workload b64 --steps 1 --warmup 1 > gpurun_out/plain_b64.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'dense_small_kernel' -s 1 -c 1 -o gpurun_out/r2_k1s_b64 \
    python tools/prof_step.py --workload b64 --steps 1 --warmup 1 > gpurun_out/ncu_k1s.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -8
