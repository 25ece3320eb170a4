#!/bin/bash
for lib in "" scratch/libhem_GATE_TB_2.so scratch/libhem_GATE_TB_8.so scratch/libhem_MOM_B_4.so scratch/libhem_MOM_B_16.so; do
  echo "#### lib=$lib"
  HEM_B200_LIB=$([ -n "$lib" ] && echo $PWD/$lib) python - <<'PY'
import sys; sys.path.insert(0,'.')
from scripts.gpu_check import timing
timing(1_000_000, 5000, 5, reps=3)
PY
done
