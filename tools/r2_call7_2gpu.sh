# 2-GPU call (charged 2x): token agreement of the fused NVLink exchange, TP vs the one-GPU model, the bench at N=2, sharded quantise-only.
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
{
timeout 200 $R --master-port 29551 tools/tp_check.py 2>&1 | grep -E "nccl|p2p|AGREE|Error|error" | tail -5
timeout 200 $R --master-port 29552 tools/tp_vs_single.py 2>&1 | grep -E "single|tp2|AGREE|Error|error" | tail -5
timeout 200 $R --master-port 29555 tools/tp_quant_check.py 2>&1 | grep -E "nbits|SHARDED|Error|error" | tail -10
} 2>&1 | tee gpurun_out/c7_tp.log
timeout 400 $R --master-port 29553 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/c7_bench_tp2.json 2> gpurun_out/c7_bench_tp2.err; head -c 900 gpurun_out/c7_bench_tp2.json; echo; tail -3 gpurun_out/c7_bench_tp2.err
timeout 200 $R --master-port 29554 tools/quantize_sharded.py --model 8b > gpurun_out/c7_quant_8b_2gpu.json 2>&1; tail -1 gpurun_out/c7_quant_8b_2gpu.json | head -c 400; echo
timeout 200 python tools/quantize_sharded.py --model 8b > gpurun_out/c7_quant_8b_1gpu.json 2>&1; tail -1 gpurun_out/c7_quant_8b_1gpu.json | head -c 400; echo
timeout 300 python -m pytest tests/test_harness_gpu.py tests/test_hqq_linear_gpu.py tests/test_zz_tp_shards_gpu.py -m gpu -q 2>&1 | tail -8 | tee gpurun_out/c7_tests.log
