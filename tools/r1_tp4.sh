mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 100 --warmup 5 > gpurun_out/bench_tp4.json 2> gpurun_out/bench_tp4.err
grep '"metric"' gpurun_out/bench_tp4.json | head -c 400
