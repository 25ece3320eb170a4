#!/bin/bash
# gate bring-up: small parity checks first (fail fast, bounded), then timings (optionally for alternative builds of the library)
mkdir -p gpurun_out
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/gate2_check.log
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'scripts')
import gpu_check as g
ok = True
for a in [(3000, 20000, 3), (20000, 5000, 4), (100000, 5000, 3), (50000, 1000, 3)]:
    ok &= g.check(*a)
print("ALL", "OK" if ok else "MISMATCH")
PY
timeout 300 python scripts/gpu_base3m.py 2>&1 | tee gpurun_out/gate2_time.log
for lib in "$@"; do
  echo "=== variant $lib"
  HEM_B200_LIB=$PWD/gaussiansplattingregistration_b200/$lib timeout 300 python scripts/gpu_base3m.py 2>&1 | tee gpurun_out/gate2_time_$lib.log
done
