set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bm25 or multipath" > gpurun_out/r1w_pytest.log 2>&1; tail -15 gpurun_out/r1w_pytest.log
timeout 600 python tools/bench_extra.py c4s > gpurun_out/r1w_c4s.json 2> gpurun_out/r1w_c4s.err; cut -c1-300 gpurun_out/r1w_c4s.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"zero_fill" -c 6 --csv --log-file gpurun_out/r1w_zf.csv python tools/bench_extra.py c4s > gpurun_out/r1w_ncu.log 2>&1; tail -3 gpurun_out/r1w_zf.csv | cut -c60-120,250-
