# two GPUs: the NCCL parity test and one sharded bench line (stdout must hold exactly the JSON line)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/m2_pytest.log 2>&1; tail -3 gpurun_out/m2_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/m2_bench.json 2> gpurun_out/m2_bench.err; wc -l gpurun_out/m2_bench.json; cut -c1-160 gpurun_out/m2_bench.json
timeout 300 python bench.py --workload small --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/m2_small.json 2> gpurun_out/m2_small.err; wc -l gpurun_out/m2_small.json; cut -c1-120 gpurun_out/m2_small.json
