# Round-2 call 1: which of the round-1 blind variants are right and which win; ncu of solver / gemm / decode winner.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/c1_smi.log
export HQQ_B200_RUN_EXPERIMENTAL=1
timeout 700 python -m pytest tests/test_zz_variants_gpu.py -m gpu -q -n 4 2>&1 | tail -25 | tee gpurun_out/c1_variants_tests.log
unset HQQ_B200_RUN_EXPERIMENTAL
{
echo "== decode step per D1 variant"
for v in 0 3042 1042 2042 7042 1033; do
  echo -n "D1_VARIANT=$v: "; HQQ_B200_D1_VARIANT=$v timeout 120 python tools/step_time.py 2>&1 | tail -1
done
for mb in 16 48; do
  echo -n "WPF_MB=$mb + 3042: "; HQQ_B200_WPF_MB=$mb HQQ_B200_D1_VARIANT=3042 timeout 120 python tools/step_time.py 2>&1 | tail -1
done
echo -n "WPF 64 o,gu + 3042: "; HQQ_B200_WPF_MB=64 HQQ_B200_WPF_AHEAD=1 HQQ_B200_WPF_FROM=o,gu HQQ_B200_D1_VARIANT=3042 timeout 120 python tools/step_time.py 2>&1 | tail -1
echo "== quantizer"
timeout 200 python tools/prof_quantize.py 8b 4
HQQ_B200_SOLVER_VARIANT=1 timeout 200 python tools/prof_quantize.py 8b 4,2
echo "== GEMM M=4096 variants"
for g in "" ld un512 dq16 un512dq ld512; do
  echo "-- GEMM_VARIANT=$g"; HQQ_B200_GEMM_VARIANT=$g timeout 200 python tools/prof_gemm.py 1024,4096 4 2>&1 | grep -E "fused|cublas"
done
echo "== 3-bit one token"
timeout 200 python tools/prof_gemm.py 1 3 2>&1 | grep fused
HQQ_B200_FUSED_3BIT=1 timeout 200 python tools/prof_gemm.py 1 3 2>&1 | grep fused
echo "== mid M default vs split-K"
timeout 200 python tools/prof_gemm.py 64,128,256,512 4 2>&1 | grep -E "fused|cublas"
HQQ_B200_GEMM_SPLITK=1 timeout 200 python tools/prof_gemm.py 64,128,256,512 4 2>&1 | grep fused
} 2>&1 | tee gpurun_out/c1_sweep.log
# ncu: fast solver + quant_pack on 14336x4096, decode 3042 (gate+up), gemm default at 4096
HQQ_B200_SOLVER_VARIANT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'solver_axis1|quant_pack' -s 4 -c 2 -o gpurun_out/c1_solver python tools/prof_one_quant.py > gpurun_out/c1_ncu_solver.log 2>&1
HQQ_B200_D1_VARIANT=3042 LAYERS=2 REPS=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_decode1 -s 40 -c 4 -o gpurun_out/c1_decode python tools/step_time.py > gpurun_out/c1_ncu_decode.log 2>&1
ls -la gpurun_out | tail -20
