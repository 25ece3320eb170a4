#!/bin/bash
for lib in "" scratch/libhem_mb5.so scratch/libhem_mb6.so; do
  echo "#### lib=$lib"
  HEM_B200_LIB=$([ -n "$lib" ] && echo $PWD/$lib) python - <<'PY'
import sys; sys.path.insert(0,'.')
from scripts.gpu_check import timing, check
check(20000, 5000, 3)
timing(1_000_000, 5000, 5, reps=3)
PY
done
