set -x
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1h_plain_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"dense_fused_pair_kernel" -s 7 -c 1 -o gpurun_out/r1h_prof_c3 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1h_ncu_c3.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1h_plain_c3b.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1h_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1h_ncu_c3b.log 2>&1
python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1h_plain_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1h_launches_c2.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1h_ncu_c2.log 2>&1
python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1h_plain_c2b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"dense_fused_pair_kernel" -s 7 -c 1 -o gpurun_out/r1h_prof_c2 -f python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1h_ncu_c2b.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r1h_bench_c3.json 2> gpurun_out/r1h_bench_c3.err
python bench.py --workload c2 --steps 50 --warmup 5 > gpurun_out/r1h_bench_c2.json 2> gpurun_out/r1h_bench_c2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1h_bench_ref.json 2> gpurun_out/r1h_bench_ref.err
cut -c1-300 gpurun_out/r1h_bench_c3.json gpurun_out/r1h_bench_c2.json gpurun_out/r1h_bench_ref.json
