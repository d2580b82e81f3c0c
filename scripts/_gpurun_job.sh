timeout 300 python scripts/align_e2e_profile.py 2>&1 | grep -v Warn | head -50
