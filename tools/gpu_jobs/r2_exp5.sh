#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q -k "exact or online or random or concurrent or certificate or ties" 2>&1 | tail -4
python tools/online_call.py > gpurun_out/r2_online_call.json 2>gpurun_out/e.err; grep -E "\"dense|device_ms|achieved|cpu_port_ms|equals" gpurun_out/r2_online_call.json | head -40
