# what the driver runs at round end on one B200: the GPU tests, smoke(), the default bench line and the reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/dl_pytest.log 2>&1; tail -3 gpurun_out/dl_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/dl_smoke.log 2>&1; tail -1 gpurun_out/dl_smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/dl_bench_ref.json 2> gpurun_out/dl_bench_ref.err; cut -c1-200 gpurun_out/dl_bench_ref.json
timeout 900 python bench.py > gpurun_out/dl_bench.json 2> gpurun_out/dl_bench.err; cat gpurun_out/dl_bench.json
