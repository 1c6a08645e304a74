# Second 8-GPU call of round 2 (charged 8x: short).  After the fixes: teacher-forced p2p / NCCL check, argmax with the key exchange
# in the launch, OpenMP pinning confined to the CPU legs (the first call's e2e values were throttled by it).
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 120 $R --nproc-per-node 8 --master-port 29621 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/n8_bench_8b_tp8.json 2> gpurun_out/n8_bench_8b_tp8.err; grep '"metric"' gpurun_out/n8_bench_8b_tp8.json | cut -c1-1500; tail -2 gpurun_out/n8_bench_8b_tp8.err | cut -c1-600
timeout 200 $R --nproc-per-node 8 --master-port 29622 bench.py --gpus 8 --steps 50 --warmup 5 --model 70b > gpurun_out/n8_bench_70b_tp8.json 2> gpurun_out/n8_bench_70b_tp8.err; grep '"metric"' gpurun_out/n8_bench_70b_tp8.json | cut -c1-1500
timeout 100 $R --nproc-per-node 4 --master-port 29623 bench.py --gpus 4 --steps 100 --warmup 5 > gpurun_out/n8_bench_8b_tp4.json 2> gpurun_out/n8_bench_8b_tp4.err; grep '"metric"' gpurun_out/n8_bench_8b_tp4.json | cut -c1-1500
timeout 200 $R --nproc-per-node 8 --master-port 29624 bench.py --gpus 8 --steps 30 --warmup 5 --model 70b --batch 32 > gpurun_out/n8_bench_70b_tp8_bs32.json 2> gpurun_out/n8_bench_70b_tp8_bs32.err; grep '"metric"' gpurun_out/n8_bench_70b_tp8_bs32.json | cut -c1-900
