set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tail_variants" > gpurun_out/r1p_pytest.log 2>&1; tail -3 gpurun_out/r1p_pytest.log
for t in 7 8 5; do
  for wl in c3s8 c2; do
  timeout 300 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline --tail $t > gpurun_out/r1p_${wl}_tail$t.json 2> gpurun_out/r1p_${wl}_tail$t.err
  python -c "import json;d=json.load(open('gpurun_out/r1p_${wl}_tail$t.json'));print('$wl tail',$t,d['ms_per_step'],d['roofline']['kernel_ms'],d['search']['tail_ms'])"
  done
done
for t in 7 8; do
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:"cand_reduce|rescore" -c 12 --csv --log-file gpurun_out/r1p_launches_c3s8_t$t.csv python bench.py --workload c3s8 --steps 2 --warmup 3 --no-cpu-baseline --tail $t > gpurun_out/r1p_ncu_c3s8.log 2>&1
tail -1 gpurun_out/r1p_launches_c3s8_t$t.csv | cut -c60-130,380-
done
