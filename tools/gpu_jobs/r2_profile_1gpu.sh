#!/bin/bash
# round 2, one GPU: tests after the kernel changes, benches, ncu launch lists and full captures of the changed kernels
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_reference_boundary.py -m gpu -x -q 2>&1 | tail -6
for wl in c2 c4s8 c3s8; do
  python bench.py --workload $wl --no-cpu-baseline 2>gpurun_out/r2_$wl.err > gpurun_out/r2_bench_${wl}_1gpu.json || tail -5 gpurun_out/r2_$wl.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench_${wl}_1gpu.json"))
    print("$wl", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"],3), "kms", round(d["roofline"]["kernel_ms"],4), "parity", {k:v["ok"] for k,v in d["parity"].items()}, d.get("search"))
    for r in d.get("rooflines", []): print("   ", r["kernel"], round(r["kernel_ms"],4), round(r["frac"],3))
except Exception as e: print("$wl failed", e)
PY
done
for wl in c2 c3s8 c4s8 online; do
  python tools/prof_step.py --workload $wl --steps 2 --warmup 2 > gpurun_out/plain_$wl.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_$wl.csv \
      python tools/prof_step.py --workload $wl --steps 2 --warmup 2 > gpurun_out/ncu_$wl.log 2>&1
done
python tools/prof_step.py --workload c2 --steps 1 --warmup 2 > gpurun_out/plain_c2b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'rescore_finalize|cand_reduce|tau_from' -s 4 -c 3 -o gpurun_out/r2_tail_c2 \
    python tools/prof_step.py --workload c2 --steps 1 --warmup 2 > gpurun_out/ncu_tail.log 2>&1
python tools/prof_step.py --workload online --steps 1 --warmup 1 > gpurun_out/plain_onl.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'exact_scores' -s 1 -c 1 -o gpurun_out/r2_exact_online \
    python tools/prof_step.py --workload online --steps 1 --warmup 1 > gpurun_out/ncu_exact.log 2>&1
python tools/prof_step.py --workload c3s8 --steps 1 --warmup 1 > gpurun_out/plain_c3s8.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'dense_fused_pair_kernel' -s 3 -c 1 -o gpurun_out/r2_k1_c3s8 \
    python tools/prof_step.py --workload c3s8 --steps 1 --warmup 1 > gpurun_out/ncu_k1.log 2>&1
ls -la gpurun_out/*.ncu-rep
