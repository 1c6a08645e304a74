# Round-2 call 4: the whole GPU suite (call 3 stopped at the removed L2-hint test), GEMM with the L2 prefetch, route 3 timings.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 -n 3 2>&1 | tail -60 > gpurun_out/c4_tests.log; tail -4 gpurun_out/c4_tests.log
{
echo "== GEMM"; timeout 300 python tools/prof_gemm.py 64,128,256,512,1024,4096 4 2>&1 | grep -E "fused|cublas"
echo "== 3-bit (route 3)"; timeout 300 python tools/prof_gemm.py 1,64,4096 3 2>&1 | grep -E "fused|cublas"
} 2>&1 | tee gpurun_out/c4_perf.log
ls -la gpurun_out | tail -5
