set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; tail -3 gpurun_out/final_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench_c3.json 2> gpurun_out/final_bench_c3.err; cut -c1-300 gpurun_out/final_bench_c3.json
timeout 600 python bench.py --workload c2 --steps 50 --warmup 5 > gpurun_out/final_bench_c2.json 2> gpurun_out/final_bench_c2.err; cut -c1-300 gpurun_out/final_bench_c2.json
timeout 600 python bench.py --workload c3s8 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/final_bench_c3s8.json 2> gpurun_out/final_bench_c3s8.err; cut -c1-300 gpurun_out/final_bench_c3s8.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu_c3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_c2.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"cand_reduce_kernel|rescore_finalize_kernel" -s 8 -c 2 -o gpurun_out/final_prof_tail -f python bench.py --workload c3s8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu_tail.log 2>&1
ls -la gpurun_out/r1q*
