#!/bin/bash
# round 2, N GPUs (gpurun --gpus N): multi-GPU parity (dense + hybrid, both exchange routes, pipelined searcher), then the
# benches at N ranks.   usage: bash tools/gpu_jobs/r2_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15
run() {  # name, extra args...
  name=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" \
      2>gpurun_out/r2_${name}_${N}gpu.err > gpurun_out/r2_bench_${name}_${N}gpu.json || tail -8 gpurun_out/r2_${name}_${N}gpu.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench_${name}_${N}gpu.json"))
    print("$name N=$N", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"],3), "kms", round(d["roofline"]["kernel_ms"],4), "parity", {k:v["ok"] for k,v in d["parity"].items()}, d.get("latency_ms"), d.get("run"))
    for r in d.get("rooflines", []): print("   ", r["kernel"], round(r["kernel_ms"],4), round(r["frac"],3))
except Exception as e: print("$name failed", e)
PY
}
run c3 --steps 100 --warmup 10
run c3_nccl --steps 100 --warmup 10 --exchange nccl
run c3_sync --steps 100 --warmup 10 --sync
run c4 --workload c4 --steps 30 --warmup 5
if [ "$N" -lt 8 ]; then run c4_nccl --workload c4 --steps 30 --warmup 5 --exchange nccl; fi
if [ "$N" -ge 8 ]; then run c5 --workload c5; fi
