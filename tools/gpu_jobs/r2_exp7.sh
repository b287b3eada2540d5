#!/bin/bash
mkdir -p gpurun_out
for p in 1 0; do
  VFI_TRACE_STEPS=1 timeout 200 python tools/trace_steps.py --workload c3s8 --profile $p 2>&1 | grep -A40 "^--- c3s8" | tail -16
done
VFI_TRACE_STEPS=1 timeout 200 python tools/trace_steps.py --workload c2 --profile 0 2>&1 | grep -A40 "^--- c2" | tail -8
