# Round-2 first call: decode step time under each experimental M = 1 kernel variant (same box, same process layout), then the
# GEMM variants.  Usage: gpurun --timeout 900 -- 'bash tools/variant_sweep.sh'
export HQQ_B200_RUN_EXPERIMENTAL=1
timeout 600 python -m pytest tests/test_zz_variants_gpu.py -m gpu -q 2>&1 | tail -5
for v in 0 1042 2042 3042 4042 7042 1033 7033; do
  HQQ_B200_D1_VARIANT=$v timeout 120 python tools/step_time.py 2>&1 | tail -1
done
timeout 200 python tools/prof_gemm.py 4096 4 2>&1 | grep fused
HQQ_B200_GEMM_VARIANT=ld timeout 200 python tools/prof_gemm.py 4096 4 2>&1 | grep fused
HQQ_B200_GEMM_VARIANT=ld HQQ_B200_GEMM_UN=128 timeout 200 python tools/prof_gemm.py 4096 4 2>&1 | grep fused
HQQ_B200_GEMM_UN=128 timeout 200 python tools/prof_gemm.py 4096 4 2>&1 | grep fused
# BASELINE configs[2]: the per-linear sweep (M x nbits x 3 shapes), kept under profiles/ afterwards
timeout 900 python tools/prof_gemm.py 1,16,32,128,1024,4096 8,4,3,2,1 > gpurun_out/gemm_sweep.log 2>&1; grep -c fused gpurun_out/gemm_sweep.log
