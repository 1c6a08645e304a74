# 1 GPU: GPU suite, split-K timing with the second pass as a programmatic dependent, launch list of a 4-block bench run, default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c17_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c17_pytest.log; tail -3 gpurun_out/c17_pytest.log
timeout 300 python tools/prof_splitk.py 4096x4096 > gpurun_out/c17_splitk.log 2>&1; cat gpurun_out/c17_splitk.log
timeout 300 python tools/prof_splitk.py 4096x14336 >> gpurun_out/c17_splitk.log 2>&1; tail -10 gpurun_out/c17_splitk.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/c17_launches_4layers.csv python bench.py --layers 4 --steps 2 --warmup 3 --no-extras > gpurun_out/c17_launch_bench.log 2>&1; wc -l gpurun_out/c17_launches_4layers.csv
timeout 600 python bench.py > gpurun_out/c17_bench.json 2> gpurun_out/c17_bench.err; head -c 600 gpurun_out/c17_bench.json; echo
