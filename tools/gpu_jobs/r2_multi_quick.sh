#!/bin/bash
# quick multi-GPU check: parity test + C3 on the NCCL route     usage: bash tools/gpu_jobs/r2_multi_quick.sh N
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 --exchange nccl \
    2>gpurun_out/r2_c3_nccl_${N}gpu.err > gpurun_out/r2_bench_c3_nccl_${N}gpu.json || grep -v "^\[W" gpurun_out/r2_c3_nccl_${N}gpu.err | tail -12
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_c3_nccl_${N}gpu.json"))
print("c3_nccl N=$N", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],3), "parity", {k:v["ok"] for k,v in d["parity"].items()}, d["run"])
PY
