timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -6
