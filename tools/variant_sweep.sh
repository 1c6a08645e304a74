# Round-2 first call: validate every experimental variant (bit-identity tests), then time each against the default on the same
# box.  Usage: gpurun --timeout 1500 -- 'bash tools/variant_sweep.sh'
mkdir -p gpurun_out
export HQQ_B200_RUN_EXPERIMENTAL=1
timeout 900 python -m pytest tests/test_zz_variants_gpu.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/variants_tests.log
unset HQQ_B200_RUN_EXPERIMENTAL
{
echo "== decode step under each M = 1 kernel variant"
for v in 0 1042 2042 3042 4042 7042 1033 7033; do
  echo -n "D1_VARIANT=$v: "; HQQ_B200_D1_VARIANT=$v timeout 120 python tools/step_time.py 2>&1 | tail -1
done
echo "== decode step with cross-launch weight prefetch (MiB per launch; > 24 covers two launches ahead)"
for mb in 4 8 16 32 48 64; do
  echo -n "WPF_MB=$mb: "; HQQ_B200_WPF_MB=$mb timeout 120 python tools/step_time.py 2>&1 | tail -1
  echo -n "WPF_MB=$mb + D1 7042: "; HQQ_B200_WPF_MB=$mb HQQ_B200_D1_VARIANT=7042 timeout 120 python tools/step_time.py 2>&1 | tail -1
done
echo "== weight prefetch only from the launches whose prologue runs while HBM idles (o: under attention, gu: under o)"
for cfg in "48 1 o" "64 1 o" "64 1 o,gu" "88 2 o" "64 1 o,gu,qkv"; do
  set -- $cfg
  echo -n "WPF_MB=$1 AHEAD=$2 FROM=$3: "; HQQ_B200_WPF_MB=$1 HQQ_B200_WPF_AHEAD=$2 HQQ_B200_WPF_FROM=$3 timeout 120 python tools/step_time.py 2>&1 | tail -1
done
echo "== quantizer: default vs fast solver (HQQ_B200_SOLVER_VARIANT=1), Llama-3-8B and 70B layer shapes"
timeout 200 python tools/prof_quantize.py 8b 4
HQQ_B200_SOLVER_VARIANT=1 timeout 200 python tools/prof_quantize.py 8b 4,2
HQQ_B200_SOLVER_VARIANT=1 timeout 200 python tools/prof_quantize.py 70b 4
echo "== GEMM M=4096: default / ld / un512 / UN caps"
timeout 200 python tools/prof_gemm.py 4096 4 2>&1 | grep fused
HQQ_B200_GEMM_VARIANT=un512 timeout 200 python tools/prof_gemm.py 1024,4096,8192 4 2>&1 | grep fused
HQQ_B200_GEMM_VARIANT=dq16 timeout 200 python tools/prof_gemm.py 256,1024,4096 4,2,8 2>&1 | grep fused
HQQ_B200_GEMM_VARIANT=un512dq timeout 200 python tools/prof_gemm.py 1024,4096,8192 4 2>&1 | grep fused
HQQ_B200_GEMM_VARIANT=ld timeout 200 python tools/prof_gemm.py 4096 4 2>&1 | grep fused
HQQ_B200_GEMM_VARIANT=ld512 timeout 200 python tools/prof_gemm.py 1024,4096,8192 4 2>&1 | grep fused
HQQ_B200_GEMM_VARIANT=ld HQQ_B200_GEMM_UN=128 timeout 200 python tools/prof_gemm.py 4096 4 2>&1 | grep fused
HQQ_B200_GEMM_UN=128 timeout 200 python tools/prof_gemm.py 4096 4 2>&1 | grep fused
echo "== 3-bit, one token: dequantise + GEMM (default) vs fused (HQQ_B200_FUSED_3BIT=1)"
timeout 200 python tools/prof_gemm.py 1 3 2>&1 | grep fused
HQQ_B200_FUSED_3BIT=1 timeout 200 python tools/prof_gemm.py 1 3 2>&1 | grep fused
echo "== GEMM mid M: default vs split-K"
timeout 200 python tools/prof_gemm.py 64,128,256,512 4 2>&1 | grep fused
HQQ_B200_GEMM_SPLITK=1 timeout 200 python tools/prof_gemm.py 64,128,256,512 4 2>&1 | grep fused
} 2>&1 | tee gpurun_out/variant_sweep.log
# BASELINE configs[2]: the per-linear sweep (M x nbits x 3 shapes), kept under profiles/ afterwards
timeout 900 python tools/prof_gemm.py 1,16,32,128,1024,4096 8,4,3,2,1 > gpurun_out/gemm_sweep.log 2>&1; grep -c fused gpurun_out/gemm_sweep.log
# the bench with its decode autotuner (child-process guard + in-process choice): config.autotune lists every candidate's verdict
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_autotune.json 2> gpurun_out/bench_autotune.err; head -c 1500 gpurun_out/bench_autotune.json; echo
