set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/r1g_multi8.log 2>&1; echo rc=$?; grep multi-gpu gpurun_out/r1g_multi8.log
for ex in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 50 --warmup 5 --exchange $ex > gpurun_out/r1g_g8_$ex.json 2> gpurun_out/r1g_g8_$ex.err; echo rc=$?; cut -c1-250 gpurun_out/r1g_g8_$ex.json; grep -v "^\*\|OMP\|^$" gpurun_out/r1g_g8_$ex.err | tail -3
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 50 --warmup 5 > gpurun_out/r1g_g4_peer.json 2> gpurun_out/r1g_g4_peer.err; cut -c1-250 gpurun_out/r1g_g4_peer.json
