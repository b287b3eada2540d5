mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pl_pytest.log 2>&1; tail -3 gpurun_out/pl_pytest.log
for wl in c2 c3s8; do
for mode in "" "--sync"; do
timeout 300 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline $mode > gpurun_out/pl_${wl}${mode}.json 2> gpurun_out/pl_${wl}${mode}.err
python -c "import json;d=json.load(open('gpurun_out/pl_${wl}${mode}.json'));print('$wl $mode',round(d['value']),d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['search'])"
done; done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/pl_c3.json 2> gpurun_out/pl_c3.err; python -c "import json;d=json.load(open('gpurun_out/pl_c3.json'));print('c3',round(d['value']),d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['e2e'])"
