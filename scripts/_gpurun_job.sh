timeout 600 python -m pytest tests/test_align_gpu.py -x -q 2>&1 | tail -2
timeout 300 python scripts/align_stream_timeline.py 8 2>&1 | tail -2
