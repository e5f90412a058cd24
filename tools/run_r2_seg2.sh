timeout 600 python tools/chain_time.py 64 40 fused one 2>&1 | tail -12
