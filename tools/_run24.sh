timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bm25 or multipath" > gpurun_out/r1x_pytest.log 2>&1; tail -3 gpurun_out/r1x_pytest.log
