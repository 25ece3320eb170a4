#!/bin/bash
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
HEM_CELL_SCALE=${CS:-0.5} python - <<'PY'
import sys; sys.path.insert(0,'.')
from scripts.gpu_check import timing, check
ok = check(3000, 20000, 3) & check(20000, 5000, 3) & check(50000, 1000, 2)
timing(1_000_000, 5000, 5, reps=3)
timing(3_000_000, 5000, 3, reps=2, aspect=(2,1,0.5))
print("ALL", "OK" if ok else "MISMATCH")
PY
