set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/final_multi8.log 2>&1; echo rc=$?; grep multi-gpu gpurun_out/final_multi8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/final_g8_peer.json 2> gpurun_out/final_g8_peer.err; echo rc=$?; cut -c1-250 gpurun_out/final_g8_peer.json; grep -v "^\*\|OMP\|^$" gpurun_out/final_g8_peer.err | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 50 --warmup 5 > gpurun_out/final_g4_peer.json 2> gpurun_out/final_g4_peer.err; cut -c1-250 gpurun_out/final_g4_peer.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/final_g2_peer.json 2> gpurun_out/final_g2_peer.err; cut -c1-250 gpurun_out/final_g2_peer.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 8 --workload c5 --steps 300 --warmup 20 > gpurun_out/final_g8_c5.json 2> gpurun_out/final_g8_c5.err; cut -c1-400 gpurun_out/final_g8_c5.json; tail -3 gpurun_out/final_g8_c5.err
