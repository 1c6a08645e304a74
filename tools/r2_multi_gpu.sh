# Round-2 8-GPU call (charged 8x: keep it short).  Usage: gpurun --gpus 8 --timeout 900 -- 'bash tools/r2_multi_gpu.sh'
# 1. Llama-3-8B-shaped TP=8 decode (what the driver's scaling run measures at N=8), 2. Llama-3-70B-shaped TP=8 decode, bs 1 and 32
# (BASELINE configs[4]), 3. quantise-only 70B-shaped, layers sharded over 8 / 4 / 2 / 1 ranks (BASELINE configs[3]),
# 4. the tensor-parallel model against the one-GPU model on the very same quantised weights.  Every command under its own timeout.
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 150 $R --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/m8_bench_8b_tp8.json 2> gpurun_out/m8_bench_8b_tp8.err; grep '"metric"' gpurun_out/m8_bench_8b_tp8.json | head -c 500; echo
timeout 150 $R --nproc-per-node 4 --master-port 29527 bench.py --gpus 4 --steps 100 --warmup 5 > gpurun_out/m8_bench_8b_tp4.json 2> gpurun_out/m8_bench_8b_tp4.err; grep '"metric"' gpurun_out/m8_bench_8b_tp4.json | head -c 300; echo
timeout 240 $R --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --steps 50 --warmup 5 --model 70b > gpurun_out/m8_bench_70b_tp8.json 2> gpurun_out/m8_bench_70b_tp8.err; grep '"metric"' gpurun_out/m8_bench_70b_tp8.json | head -c 500; echo
timeout 240 $R --nproc-per-node 8 --master-port 29524 bench.py --gpus 8 --steps 30 --warmup 5 --model 70b --batch 32 > gpurun_out/m8_bench_70b_tp8_bs32.json 2> gpurun_out/m8_bench_70b_tp8_bs32.err; grep '"metric"' gpurun_out/m8_bench_70b_tp8_bs32.json | head -c 500; echo
for n in 8 4 2; do
  timeout 150 $R --nproc-per-node $n --master-port 2953$n tools/quantize_sharded.py --model 70b --blocks 16 > gpurun_out/m8_quant_70b_${n}gpu.json 2> gpurun_out/m8_quant_70b_${n}gpu.err; tail -1 gpurun_out/m8_quant_70b_${n}gpu.json | head -c 400; echo
done
timeout 150 python tools/quantize_sharded.py --model 70b --blocks 16 > gpurun_out/m8_quant_70b_1gpu.json 2> gpurun_out/m8_quant_70b_1gpu.err; tail -1 gpurun_out/m8_quant_70b_1gpu.json | head -c 400; echo
timeout 200 $R --nproc-per-node 8 --master-port 29526 tools/tp_vs_single.py > gpurun_out/m8_tp_vs_single.log 2>&1; grep -E "AGREE|single|tp8" gpurun_out/m8_tp_vs_single.log | head -3
