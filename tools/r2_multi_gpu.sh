# Round-2 8-GPU call (charged 8x: keep it short).  Usage: gpurun --gpus 8 --timeout 600 -- 'bash tools/r2_multi_gpu.sh'
# 1. Llama-3-8B-shaped TP=8 decode (the driver's scaling run at N=8), 2. Llama-3-70B-shaped TP=8 decode (BASELINE configs[4], bs=1),
# 3. quantise-only 70B-shaped, layers sharded over 8 ranks (BASELINE configs[3]).  Every multi-rank command under its own timeout.
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 150 $R --master-port 29521 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/bench_tp8.json 2> gpurun_out/bench_tp8.err; grep '"metric"' gpurun_out/bench_tp8.json | head -c 400; echo
timeout 200 $R --master-port 29522 bench.py --gpus 8 --steps 50 --warmup 5 --model 70b > gpurun_out/bench_70b_tp8.json 2> gpurun_out/bench_70b_tp8.err; grep '"metric"' gpurun_out/bench_70b_tp8.json | head -c 400; echo
timeout 200 $R --master-port 29524 bench.py --gpus 8 --steps 30 --warmup 5 --model 70b --batch 32 > gpurun_out/bench_70b_tp8_bs32.json 2> gpurun_out/bench_70b_tp8_bs32.err; grep '"metric"' gpurun_out/bench_70b_tp8_bs32.json | head -c 400; echo
timeout 150 $R --master-port 29523 tools/quantize_sharded.py --model 70b --blocks 16 > gpurun_out/quant_70b_8gpu.json 2> gpurun_out/quant_70b_8gpu.err; tail -1 gpurun_out/quant_70b_8gpu.json | head -c 500; echo
HQQ_B200_SOLVER_VARIANT=1 timeout 150 $R --master-port 29525 tools/quantize_sharded.py --model 70b --blocks 16 > gpurun_out/quant_70b_8gpu_fast.json 2> gpurun_out/quant_70b_8gpu_fast.err; tail -1 gpurun_out/quant_70b_8gpu_fast.json | head -c 500; echo
# the tensor-parallel model against the one-GPU model on the very same quantised weights (shards cut out of the unsharded quantisation)
timeout 200 $R --master-port 29526 tools/tp_vs_single.py > gpurun_out/tp_vs_single.log 2>&1; grep -E "AGREE|single|tp8" gpurun_out/tp_vs_single.log | head -3
