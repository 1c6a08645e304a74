# Round-2 call 3: full GPU suite on the cleaned-up build (ring M=1 kernel default, register solver default, persistent tcgen05 GEMM),
# GEMM timings, a bench run, ncu of the new GEMM kernel and of the shipped decode kernel (traffic).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -x -n 3 2>&1 | tail -40 > gpurun_out/c3_tests.log; tail -5 gpurun_out/c3_tests.log
HQQ_B200_GEMM_CTAS=7 timeout 300 python -m pytest tests/test_linear_gpu.py -m gpu -q 2>&1 | tail -5 > gpurun_out/c3_tests_ctas7.log; tail -2 gpurun_out/c3_tests_ctas7.log
{
echo "== GEMM"; timeout 300 python tools/prof_gemm.py 64,128,256,512,1024,4096 4 2>&1 | grep -E "fused|cublas"
timeout 200 python tools/prof_gemm.py 4096 8,2,1 2>&1 | grep -E "fused|cublas"
echo "== decode step"; timeout 120 python tools/step_time.py 2>&1 | tail -1
echo "== quantizer"; timeout 200 python tools/prof_quantize.py 8b 4
} 2>&1 | tee gpurun_out/c3_perf.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err; head -c 600 gpurun_out/c3_bench.json; echo; tail -3 gpurun_out/c3_bench.err
timeout 200 python bench.py --impl reference --steps 100 --warmup 5 > gpurun_out/c3_bench_ref.json 2>&1; head -c 300 gpurun_out/c3_bench_ref.json; echo
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_gemm -s 2 -c 1 -o gpurun_out/c3_gemm python tools/prof_gemm.py 4096 4 > gpurun_out/c3_ncu_gemm.log 2>&1
LAYERS=2 REPS=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_decode1 -s 40 -c 4 -o gpurun_out/c3_decode python tools/step_time.py > gpurun_out/c3_ncu_decode.log 2>&1
ls -la gpurun_out | tail -12
