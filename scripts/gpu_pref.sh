#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "prefilter or oracle_free" 2>&1 | tail -5
python - <<'PY'
import sys, time; sys.path.insert(0,'.')
import torch
from gaussiansplattingregistration_b200.synth import synth
from gaussiansplattingregistration_b200.hem import HemContext
c = synth(1_000_000, seed=1, density=5000)
dev = [torch.from_numpy(a).cuda() for a in c.arrays()]
for on in (1, 0):
    ctx = HemContext(0, 7); ctx.set_option("timing", 1); ctx.set_option("smd_prefilter", on)
    for r in range(3): ctx.build(*dev, cluster_level=5)
    for l in range(5):
        s = ctx.step_stats(l)
        print(f"prefilter {on} step {l}: np {s['n_parents']} cert {s['n_certified_parents']} gate {s['ms_gate']:.3f} k_gate {s['ms_k_gate']:.3f} k_mom {s['ms_k_moments']:.3f} tot {s['ms_total']:.3f}")
PY
