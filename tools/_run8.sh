set -x
mkdir -p gpurun_out
timeout 120 tools/ubench/fp64_rates > gpurun_out/r1i_fp64_rates.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1i_pytest.log 2>&1; tail -5 gpurun_out/r1i_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r1i_smoke.log 2>&1; tail -2 gpurun_out/r1i_smoke.log
timeout 600 python bench.py > gpurun_out/r1i_bench_default.json 2> gpurun_out/r1i_bench_default.err; cut -c1-400 gpurun_out/r1i_bench_default.json
cat gpurun_out/r1i_fp64_rates.log
