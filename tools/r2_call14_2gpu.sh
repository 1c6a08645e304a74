# 2 GPUs: the p2p / NCCL comparison on per-rank shapes of larger tensor-parallel degrees, the 2-GPU tests, the N = 2 bench line, the
# tensor-parallel model against the one-GPU model; and (one GPU idle) where the small-M kernel should hand over to the tcgen05 kernel
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $R --nproc-per-node 2 --master-port 29541 tools/tp_modes_check.py > gpurun_out/c14_tp_modes.log 2>&1; grep -E "^\{|TP_MODES" gpurun_out/c14_tp_modes.log | cut -c1-400
timeout 400 python -m pytest tests/test_harness_gpu.py tests/test_hqq_linear_gpu.py tests/test_zz_tp_shards_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/c14_pytest2.log 2>&1; tail -4 gpurun_out/c14_pytest2.log
timeout 300 $R --nproc-per-node 2 --master-port 29542 bench.py --gpus 2 --steps 200 --warmup 5 --no-extras > gpurun_out/c14_bench_tp2.json 2> gpurun_out/c14_bench_tp2.err; grep '"metric"' gpurun_out/c14_bench_tp2.json | cut -c1-1200
timeout 200 $R --nproc-per-node 2 --master-port 29543 tools/tp_vs_single.py > gpurun_out/c14_tp_vs_single.log 2>&1; grep -E "AGREE" gpurun_out/c14_tp_vs_single.log
timeout 200 $R --nproc-per-node 2 --master-port 29544 tools/tp_check.py > gpurun_out/c14_tp_check.log 2>&1; tail -3 gpurun_out/c14_tp_check.log
timeout 300 python tools/prof_route_boundary.py > gpurun_out/c14_route_boundary.log 2>&1; cat gpurun_out/c14_route_boundary.log
