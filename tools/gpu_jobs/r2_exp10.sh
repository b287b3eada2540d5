#!/bin/bash
# two GPUs: multi-GPU parity with the fused push, then A/B of the fused push against the unfused exchange kernel on 1.25M-row shards
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15
run() {  # name, env, extra args...
  name=$1; shift; envs=$1; shift
  env $envs python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" \
      2>gpurun_out/r2x_${name}_${N}gpu.err > gpurun_out/r2x_${name}_${N}gpu.json || tail -8 gpurun_out/r2x_${name}_${N}gpu.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2x_${name}_${N}gpu.json"))
    print("$name N=$N", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), "e2e_ms", round(d["e2e"]["ms_per_step"],4), "kms", round(d["roofline"]["kernel_ms"],4), "tail", round(d["search"]["tail_ms"],4), "parity", {k:v["ok"] for k,v in d["parity"].items()}, d.get("run",{}).get("re_exchanges"))
except Exception as e: print("$name failed", e)
PY
}
run c3q_fused VFI_FUSED_PUSH=1 --workload c3q --steps 200 --warmup 20 --no-cpu-baseline
run c3q_unfused VFI_FUSED_PUSH=0 --workload c3q --steps 200 --warmup 20 --no-cpu-baseline
run c3q_fused2 VFI_FUSED_PUSH=1 --workload c3q --steps 200 --warmup 20 --no-cpu-baseline
run c3q_unfused2 VFI_FUSED_PUSH=0 --workload c3q --steps 200 --warmup 20 --no-cpu-baseline
run c3q_nccl VFI_FUSED_PUSH=1 --workload c3q --steps 200 --warmup 20 --no-cpu-baseline --exchange nccl
