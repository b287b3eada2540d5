#!/bin/bash
# after the tail trimming (warp-per-query tau kernel, radix-select candidate reduction, flag count without copy operations)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
VFI_TRACE_STEPS=1 timeout 200 python tools/trace_steps.py --workload c3s8 --profile 0 2>&1 | grep -A40 "^--- c3s8" | tail -6
VFI_TRACE_STEPS=1 timeout 200 python tools/trace_steps.py --workload c2 --profile 0 2>&1 | grep -A40 "^--- c2" | tail -5
VFI_TRACE_STEPS=1 timeout 200 python tools/trace_steps.py --workload c4s8 --profile 0 2>&1 | grep -A40 "^--- c4s8" | tail -5
python bench.py --workload c3s8 --no-cpu-baseline > gpurun_out/r2b_c3s8.json 2> gpurun_out/r2b_c3s8.err; cut -c1-400 gpurun_out/r2b_c3s8.json
python bench.py --workload c2 --no-cpu-baseline > gpurun_out/r2b_c2.json 2> gpurun_out/r2b_c2.err; cut -c1-400 gpurun_out/r2b_c2.json
