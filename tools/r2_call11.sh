# 1 GPU: the whole GPU suite on the current tree (split-K GEMM, argmax with the key exchange), the mid-M table, the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c11_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c11_pytest.log
tail -5 gpurun_out/c11_pytest.log
timeout 300 python tools/prof_midm.py > gpurun_out/c11_midm.log 2>&1; tail -16 gpurun_out/c11_midm.log
timeout 600 python bench.py > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err; tail -c 1500 gpurun_out/c11_bench.json
