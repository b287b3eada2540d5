mkdir -p gpurun_out
for wl in c3s8 c2 b16 b64 b128; do
python bench.py --workload $wl --no-cpu-baseline 2>gpurun_out/e.err > gpurun_out/r2_bench_${wl}_1gpu.json || tail -5 gpurun_out/e.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_${wl}_1gpu.json"))
print("$wl", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"].get("pipelined_value") and round(d["e2e"]["pipelined_value"]), "ms", round(d["ms_per_step"],4), round(d["roofline"]["frac"],3), {k:v["ok"] for k,v in d["parity"].items()})
PY
done
