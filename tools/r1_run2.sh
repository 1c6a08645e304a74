timeout 600 python -m pytest tests/test_harness_gpu.py tests/test_linear_gpu.py -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/step_breakdown.py 2>&1 | tail -24
