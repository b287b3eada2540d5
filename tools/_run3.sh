set -x
python bench.py --workload c3s8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1d_plain_c3s8.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"select_rescore" -s 3 -c 1 -o gpurun_out/r1d_prof_tail -f python bench.py --workload c3s8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1d_ncu_tail.log 2>&1
tail -2 gpurun_out/r1d_ncu_tail.log
