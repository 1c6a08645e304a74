mkdir -p gpurun_out
{
echo "== quantizer"; timeout 200 python tools/prof_quantize.py 8b 4,2
echo "== bs=32, 4 layers (BASELINE configs[4] bs leg on one GPU, sanity)"; timeout 300 python bench.py --batch 32 --steps 20 --warmup 3 --layers 4 --no-extras 2>&1 | tail -1 | head -c 600; echo
} 2>&1 | tee gpurun_out/c8_perf.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -n 3 2>&1 | tail -30 > gpurun_out/c8_tests.log; tail -3 gpurun_out/c8_tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'solver_axis1|quant_pack' -s 4 -c 2 -o gpurun_out/c8_solver python tools/prof_one_quant.py > gpurun_out/c8_ncu_solver.log 2>&1
