# 1 GPU: the whole GPU suite, the default bench line (with the nested objects), the launch list of the bench, one ncu capture of the
# split-K pair (tcgen05 kernel + second pass) at M = 128 on 4096 x 4096
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c16_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c16_pytest.log; tail -4 gpurun_out/c16_pytest.log
timeout 600 python bench.py > gpurun_out/c16_bench.json 2> gpurun_out/c16_bench.err; head -c 1400 gpurun_out/c16_bench.json; echo
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"linear_gemm|splitk" -s 6 -c 2 -o gpurun_out/c16_splitk_m128 python tools/prof_gemm.py 128 4 > gpurun_out/c16_ncu.log 2>&1; tail -3 gpurun_out/c16_ncu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/c16_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/c16_launch_bench.log 2>&1; wc -l gpurun_out/c16_launches.csv
