set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size" --durations=5 > gpurun_out/r1s_pytest.log 2>&1; tail -12 gpurun_out/r1s_pytest.log
