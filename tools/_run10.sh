set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1k_pytest.log 2>&1; tail -3 gpurun_out/r1k_pytest.log
for t in 0 2; do
  for wl in c3s8 c2; do
  timeout 300 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline --tail $t > gpurun_out/r1k_${wl}_tail$t.json 2> gpurun_out/r1k_${wl}_tail$t.err
  python -c "import json;d=json.load(open('gpurun_out/r1k_${wl}_tail$t.json'));print('$wl tail',$t,d['ms_per_step'],d['roofline']['kernel_ms'],d['search']['tail_ms'])"
  done
done
ncu --set full --clock-control none --import-source on -k regex:"cand_reduce_kernel|rescore_finalize_kernel" -s 8 -c 2 -o gpurun_out/r1k_prof_tail -f python bench.py --workload c3s8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1k_ncu_tail.log 2>&1
ls -la gpurun_out/*.ncu-rep
