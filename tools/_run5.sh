set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/r1f_multi.log 2>&1; echo rc=$?; tail -8 gpurun_out/r1f_multi.log
for ex in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --exchange $ex > gpurun_out/r1f_g2_$ex.json 2> gpurun_out/r1f_g2_$ex.err; echo rc=$?; cut -c1-600 gpurun_out/r1f_g2_$ex.json; tail -3 gpurun_out/r1f_g2_$ex.err
done
