# Third 8-GPU call of round 2 (charged 8x): the bs = 32 leg of configs[4] on the batched glue kernels
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $R --nproc-per-node 8 --master-port 29631 bench.py --gpus 8 --steps 30 --warmup 5 --model 70b --batch 32 > gpurun_out/n8b_bench_70b_tp8_bs32.json 2> gpurun_out/n8b_bench_70b_tp8_bs32.err; grep '"metric"' gpurun_out/n8b_bench_70b_tp8_bs32.json | cut -c1-1200; tail -3 gpurun_out/n8b_bench_70b_tp8_bs32.err | cut -c1-500
