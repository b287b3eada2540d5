set -x
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_plain_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"dense_fused_pair_kernel|select_rescore" -s 6 -c 2 -o gpurun_out/r1c_prof_c3 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_ncu_c3.log 2>&1
tail -2 gpurun_out/r1c_plain_c3.log
tail -5 gpurun_out/r1c_ncu_c3.log
