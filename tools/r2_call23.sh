# 1 GPU, final tree: GPU suite, ncu capture of the persistent GEMM at 4096^3 (the kernel changed since r2_gemm_4096.txt: split-K code,
# programmatic dependent launch), default bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c23_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c23_pytest.log; tail -3 gpurun_out/c23_pytest.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_gemm -s 2 -c 1 -o gpurun_out/c23_gemm_4096 python tools/prof_gemm.py 4096 4 > gpurun_out/c23_ncu.log 2>&1; tail -2 gpurun_out/c23_ncu.log
timeout 600 python bench.py > gpurun_out/c23_bench.json 2> gpurun_out/c23_bench.err; head -c 700 gpurun_out/c23_bench.json; echo
