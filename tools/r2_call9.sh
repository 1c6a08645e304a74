mkdir -p gpurun_out
timeout 400 python tools/prof_midm.py 2>&1 | tee gpurun_out/c9_midm.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:linear_gemm -s 2 -c 1 -o gpurun_out/c9_gemm_m128 python tools/prof_gemm.py 128 4 > gpurun_out/c9_ncu.log 2>&1
