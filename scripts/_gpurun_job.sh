mkdir -p gpurun_out
timeout 1700 python -m pytest tests/test_forward_gpu.py -q -s -k "batched_forward_path or many_ar" 2>&1 | grep -v Warn | grep -E "Error|error|assert|rel-L2|worst|passed|failed|^E " | head -40
