set -x
timeout 600 python -m pytest tests/test_align_gpu.py -x -q 2>&1 | tail -15
timeout 600 python scripts/align_config5.py c3 n24 c5 > gpurun_out/r02_align_config5_v2.jsonl 2> gpurun_out/r02_align_config5_v2.err
cat gpurun_out/r02_align_config5_v2.jsonl; tail -5 gpurun_out/r02_align_config5_v2.err
