set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bm25_zero_fill_kernel|bm25_kernel" -s 4 -c 2 -o gpurun_out/r1v_prof_bm25 -f python tools/bench_extra.py c4s > gpurun_out/r1v_ncu_bm25.log 2>&1
ls -la gpurun_out/r1v_prof_bm25.ncu-rep
