#!/bin/bash
# A/B of an environment switch on the two pyramids: bash scripts/gpu_ab.sh VAR v1 v2 ...
var=$1; shift
for v in "$@"; do
  echo "=== $var=$v"
  env $var=$v timeout 300 python scripts/gpu_base3m.py 2>&1 | grep -E "step|pyramid" | sed -E 's/.*n_in ([0-9]+).*ms tot ([0-9.]+) prep ([0-9.]+) grid ([0-9.]+) gate ([0-9.]+) \(k ([0-9.]+)\) sumlw ([0-9.]+) mom ([0-9.]+) \(k ([0-9.]+)\) comp ([0-9.]+).*/n \1 tot \2 prep \3 grid \4 gate \5 k_gate \6 sumlw \7 mom \8 k_mom \9 comp \10/'
done
