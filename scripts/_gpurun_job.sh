timeout 600 python scripts/gemm_policy_ab.py 2>/dev/null
