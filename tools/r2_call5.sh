mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 -n 3 2>&1 | tail -40 > gpurun_out/c5_tests.log; tail -4 gpurun_out/c5_tests.log
{
echo "== GEMM"; timeout 300 python tools/prof_gemm.py 64,128,256,512,1024,4096 4 2>&1 | grep -E "fused"
timeout 300 python tools/prof_gemm.py 128,4096 8,2,1 2>&1 | grep -E "fused"
} 2>&1 | tee gpurun_out/c5_perf.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_gemm -s 2 -c 1 -o gpurun_out/c5_gemm python tools/prof_gemm.py 4096 4 > gpurun_out/c5_ncu_gemm.log 2>&1
