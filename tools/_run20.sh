set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multipath or bm25 or rrf" > gpurun_out/r1t_pytest.log 2>&1; tail -3 gpurun_out/r1t_pytest.log
timeout 600 python tools/bench_extra.py c4s > gpurun_out/r1t_c4s.json 2> gpurun_out/r1t_c4s.err; cat gpurun_out/r1t_c4s.json; tail -3 gpurun_out/r1t_c4s.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r1t_launches_c4s.csv python tools/bench_extra.py c4s > gpurun_out/r1t_ncu_c4s.log 2>&1
tail -40 gpurun_out/r1t_launches_c4s.csv | cut -c60-140,300-
