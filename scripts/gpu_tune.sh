#!/bin/bash
# parity smoke + stage timings for a few grid cell scales
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for cs in 1.0 0.7 0.5 0.35; do
  echo "#### HEM_CELL_SCALE=$cs"
  HEM_CELL_SCALE=$cs python - <<'PY'
import sys; sys.path.insert(0,'.')
from scripts.gpu_check import timing, check
check(20000, 5000, 3)
timing(1_000_000, 5000, 5, reps=3)
PY
done
