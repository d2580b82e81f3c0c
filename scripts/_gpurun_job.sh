mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_align_gpu.py -x -q 2>&1 | tail -3
timeout 300 python scripts/align_stream_timeline.py 8 > gpurun_out/r02_align_stream_timeline_v4.jsonl 2>/dev/null; cat gpurun_out/r02_align_stream_timeline_v4.jsonl
timeout 600 python scripts/align_config5.py c3 n24 c5 > gpurun_out/r02_align_config5_v6.jsonl 2>/dev/null; cat gpurun_out/r02_align_config5_v6.jsonl
