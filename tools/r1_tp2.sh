mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_harness_gpu.py -m gpu -q -k tensor_parallel 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_tp2.json 2> gpurun_out/bench_tp2.err
grep '"metric"' gpurun_out/bench_tp2.json | head -c 900
