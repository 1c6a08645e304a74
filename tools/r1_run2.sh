timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python tools/step_breakdown.py 2>&1 | head -8
