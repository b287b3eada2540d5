timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for m in 0 1 2; do
  VFI_CANON_CONV=$m python bench.py --workload c3s8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r1e_plain_$m.log 2>&1 && python - <<PY
import json
d=json.loads(open("gpurun_out/r1e_plain_$m.log").read().strip().splitlines()[-1])
print("conv $m step", round(d["ms_per_step"],3), "kernel", round(d["roofline"]["kernel_ms"],3), "retried", d["search"]["retried_queries"])
PY
done
for m in 0 1 2; do
  VFI_CANON_CONV=$m ncu --metrics gpu__time_duration.sum --clock-control none -k regex:select_rescore -s 3 -c 2 --csv --log-file gpurun_out/r1e_tail_$m.csv python bench.py --workload c3s8 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  tail -2 gpurun_out/r1e_tail_$m.csv | cut -d, -f5,13-
done
