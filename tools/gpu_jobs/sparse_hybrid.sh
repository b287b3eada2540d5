timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bm25 or multipath" > gpurun_out/r1y_pytest.log 2>&1; tail -3 gpurun_out/r1y_pytest.log
timeout 600 python tools/bench_extra.py bm25 > gpurun_out/r1y_bm25.json 2> gpurun_out/r1y_bm25.err; cat gpurun_out/r1y_bm25.json
timeout 600 python tools/bench_extra.py c4s > gpurun_out/r1y_c4s.json 2> gpurun_out/r1y_c4s.err; cat gpurun_out/r1y_c4s.json
