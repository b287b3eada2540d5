set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1l_pytest.log 2>&1; tail -3 gpurun_out/r1l_pytest.log
for t in 0 2; do
  for wl in c3s8 c2; do
  timeout 300 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline --tail $t > gpurun_out/r1l_${wl}_tail$t.json 2> gpurun_out/r1l_${wl}_tail$t.err
  python -c "import json;d=json.load(open('gpurun_out/r1l_${wl}_tail$t.json'));print('$wl tail',$t,d['ms_per_step'],d['roofline']['kernel_ms'],d['search']['tail_ms'])"
  done
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1l_launches_c3s8.csv python bench.py --workload c3s8 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1l_ncu_c3s8.log 2>&1
tail -6 gpurun_out/r1l_launches_c3s8.csv | cut -c60-120,400-
