VFI_BENCH_BREAKDOWN=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1z_c3.json 2> gpurun_out/r1z_c3.err; grep breakdown gpurun_out/r1z_c3.err; python -c "
import json; d=json.load(open('gpurun_out/r1z_c3.json')); print(d['ms_per_step'], d['e2e'])"
VFI_BENCH_BREAKDOWN=1 timeout 600 python bench.py --workload c2 --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/r1z_c2.json 2> gpurun_out/r1z_c2.err; grep breakdown gpurun_out/r1z_c2.err; python -c "
import json; d=json.load(open('gpurun_out/r1z_c2.json')); print(d['ms_per_step'], d['e2e'])"
python - <<'PY'
import torch, time
q=torch.empty((1024,1024),dtype=torch.float32).pin_memory(); d=torch.empty((1024,1024),dtype=torch.float32,device='cuda')
o=torch.empty((1024,100),dtype=torch.int64).pin_memory(); od=torch.zeros((1024,100),dtype=torch.int64,device='cuda')
for _ in range(3): d.copy_(q,non_blocking=True); o.copy_(od,non_blocking=True)
torch.cuda.synchronize()
e0,e1,e2=[torch.cuda.Event(enable_timing=True) for _ in range(3)]
e0.record(); 
for _ in range(20): d.copy_(q,non_blocking=True)
e1.record()
for _ in range(20): o.copy_(od,non_blocking=True)
e2.record(); torch.cuda.synchronize()
print('h2d 4MB ms', e0.elapsed_time(e1)/20, 'd2h 0.8MB ms', e1.elapsed_time(e2)/20)
PY
