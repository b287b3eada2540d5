set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1j_pytest.log 2>&1; tail -5 gpurun_out/r1j_pytest.log
for t in 0 1 2; do
  timeout 300 python bench.py --workload c3s8 --steps 50 --warmup 5 --no-cpu-baseline --tail $t > gpurun_out/r1j_c3s8_tail$t.json 2> gpurun_out/r1j_c3s8_tail$t.err
  python -c "import json;d=json.load(open('gpurun_out/r1j_c3s8_tail$t.json'));print('c3s8 tail',$t,d['ms_per_step'],d['roofline']['kernel_ms'],d['search'])"
  timeout 300 python bench.py --workload c2 --steps 50 --warmup 5 --no-cpu-baseline --tail $t > gpurun_out/r1j_c2_tail$t.json 2> gpurun_out/r1j_c2_tail$t.err
  python -c "import json;d=json.load(open('gpurun_out/r1j_c2_tail$t.json'));print('c2 tail',$t,d['ms_per_step'],d['roofline']['kernel_ms'],d['search'])"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1j_launches_c3s8.csv python bench.py --workload c3s8 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1j_ncu_c3s8.log 2>&1
tail -12 gpurun_out/r1j_launches_c3s8.csv | cut -c1-200
