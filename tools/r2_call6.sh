mkdir -p gpurun_out
{
echo "== GEMM smoke"; timeout 120 python tools/prof_gemm.py 128,4096 4 2>&1 | grep -E "fused"
echo "== chain check"; timeout 120 python tools/chain_check.py 2>&1 | tail -2
echo "== decode step, chain on / off"; timeout 150 python tools/step_time.py 2>&1 | tail -1; HQQ_B200_DECODE_CHAIN=0 timeout 150 python tools/step_time.py 2>&1 | tail -1
} 2>&1 | tee gpurun_out/c6_perf.log
timeout 600 python -m pytest tests -m gpu -q --maxfail=30 -n 3 2>&1 | tail -30 > gpurun_out/c6_tests.log; tail -3 gpurun_out/c6_tests.log
timeout 200 python tools/prof_gemm.py 64,256,512,1024 4 2>&1 | grep -E "fused" | tee -a gpurun_out/c6_perf.log
