mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense_search or ties or incremental" > gpurun_out/r2b_pytest.log 2>&1; tail -3 gpurun_out/r2b_pytest.log
timeout 300 python tools/bench_extra.py c5s > gpurun_out/c5s.json 2>gpurun_out/c5s.err; cat gpurun_out/c5s.json
