#!/bin/bash
# time alternative builds of the library (libhem_v_*.so) on the 3M / 1M pyramids: level-0/1 kernel times
mkdir -p gpurun_out
for lib in libhem_b200.so "$@"; do
  echo "=== $lib"
  HEM_B200_LIB=$PWD/gaussiansplattingregistration_b200/$lib timeout 300 python scripts/gpu_base3m.py 2>&1 | grep -E "step [01]:|pyramid" | sed -E 's/.*n_in ([0-9]+).*gate ([0-9.]+) \(k ([0-9.]+)\).*mom ([0-9.]+) \(k ([0-9.]+)\).*/n \1 gate \2 k_gate \3 mom \4 k_mom \5/' | tee -a gpurun_out/variants.log
done
