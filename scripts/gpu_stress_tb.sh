#!/bin/bash
# ordinary bench + stress cloud against the row threshold of the heavy-parent gate split (floor / relative factor)
for T in "256 8"; do set -- $T; echo "== HEM_HEAVY_ROWS=$1 HEM_HEAVY_REL=$2"
  HEM_HEAVY_ROWS=$1 HEM_HEAVY_REL=$2 bash scripts/gpu_stress.sh 2>&1 | grep "build 3" | cut -c1-200
  HEM_HEAVY_ROWS=$1 HEM_HEAVY_REL=$2 timeout 200 python bench.py --skip-e2e --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | cut -c1-160
done
