timeout 300 python scripts/gpu_base3m.py 2>&1 | grep -E "step|pyramid" | cut -c1-260
