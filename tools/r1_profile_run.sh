set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_harness_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 200 ncu --set full --clock-control none -k regex:linear_decode1 -f -o gpurun_out/prof_r1_groups python tools/prof_groups.py > gpurun_out/prof_groups.log 2>&1
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 2 --warmup 3 --layers 4 > gpurun_out/ncu_bench.log 2>&1
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 1500 gpurun_out/bench.json
