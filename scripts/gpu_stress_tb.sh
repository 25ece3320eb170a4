#!/bin/bash
for T in 8 4 2 1; do echo "== HEM_GATE_TICKET=$T"; HEM_GATE_TICKET=$T bash scripts/gpu_stress.sh 2>&1 | grep "build 3" | cut -c1-200; done
