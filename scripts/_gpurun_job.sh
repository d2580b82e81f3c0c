timeout 600 python -m pytest tests/test_forward_gpu.py -x -q -k "attention" 2>&1 | tail -3
timeout 300 python scripts/attn_bench.py 2>&1 | grep '"impl": 3,\|diff'
