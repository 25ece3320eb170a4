#!/bin/bash
for cs in 0.4 0.5 0.65 0.8; do
  echo "#### HEM_CELL_SCALE=$cs"
  HEM_CELL_SCALE=$cs python - <<'PY'
import sys; sys.path.insert(0,'.')
from scripts.gpu_check import timing
timing(1_000_000, 5000, 5, reps=3)
timing(3_000_000, 5000, 3, reps=2, aspect=(2,1,0.5))
PY
done
