#!/bin/bash
# at the final HEAD: the driver's sequence once more (tests, smoke, default bench), the launch lists and the step trace
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 2>gpurun_out/e.err > gpurun_out/r2_bench_c3_1gpu_20steps.json || tail -5 gpurun_out/e.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_c3_1gpu_20steps.json"))
print("c3 20 steps", round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"],3), d["parity"]["small"]["ok"], d["parity"]["timed"]["ok"])
PY
for wl in c3 c2 c4s8; do
  python tools/prof_step.py --workload $wl --steps 2 --warmup 2 > gpurun_out/plain_$wl.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_$wl.csv \
      python tools/prof_step.py --workload $wl --steps 2 --warmup 2 > gpurun_out/ncu_$wl.log 2>&1
done
: > gpurun_out/r2_step_trace.txt
for wl in c3s8 c2 c3; do
  VFI_TRACE_STEPS=1 python tools/trace_steps.py --workload $wl --profile 0 --steps 8 2>&1 | grep -A12 "^--- $wl" | tail -6 | sed "s/^/[$wl] /" >> gpurun_out/r2_step_trace.txt
done
tail -4 gpurun_out/r2_step_trace.txt
