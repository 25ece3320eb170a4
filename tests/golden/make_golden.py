"""Generate the golden fixtures in this directory from the REAL reference.

Runs only where /root/reference exists and `make -C oracle ref` has produced
oracle/_ref/libhemref.so (the unmodified reference cpp_ext + white-box driver).
Each .npz holds the inputs, the parameters, every level the reference produced
(including the weights and parent flags the reference API does not export) and
the per-parent child lists, so that the oracle restatement and the CUDA path
can be checked on machines where the reference cannot be built.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gaussiansplattingregistration_b200.synth import synth  # noqa: E402
from oracle import refhem  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (n, density, sh_degree, levels, params(red, dd, cd, decay), seed, synth kwargs, degenerate)
    "dense_sh3": (1500, 20000.0, 3, 3, (3.0, 3.0, 2.5, 1.0), 11, {}, False),
    "sparse_sh1_params": (2500, 1000.0, 1, 2, (2.0, 2.5, 2.0, 0.7), 12, {}, False),
    "degenerate_sh0": (1200, 8000.0, 0, 2, (3.0, 3.0, 2.5, 1.0), 13, {"clip_sigma": None}, True),
}


def make(name):
    n, dens, deg, levels, (red, dd, cd, decay), seed, kw, degenerate = CASES[name]
    c = synth(n, seed=seed, density=dens, sh_degree=deg, **kw)
    if degenerate:
        # rows the reference mishandles by its own admission (README.md:114): zero, indefinite
        # and NaN covariances, exercising the NaN-passes-gates and validity-filter semantics
        rng = np.random.default_rng(99)
        idx = rng.choice(n, 60, replace=False)
        c.cov6[idx[:20]] = 0.0
        c.cov6[idx[20:40], 0] *= -1.0
        c.cov6[idx[40:50], 3] = np.nan
        c.xyz[idx[50:55], 1] = np.nan
    sh = c.sh if c.sh.shape[1] else np.zeros((n, 1), np.float32)   # reference needs features[0] (mixture.cpp:296)
    r = refhem.run(c.xyz, c.colors, c.opacity, c.cov6, sh, levels, red, dd, cd, decay, seed=seed)
    out = dict(n=n, levels=levels, params=np.array([red, dd, cd, decay], np.float32), seed=seed,
               sh_floats=c.sh.shape[1], xyz=c.xyz, colors=c.colors, opacity=c.opacity, cov6=c.cov6, sh=c.sh)
    for l in range(levels + 1):
        L = r.level(l)
        for k, v in L.items():
            if k == "sh" and c.sh.shape[1] == 0:
                v = v[:, :0]
            out[f"L{l}_{k}"] = v
    for l in range(levels):
        pidx, off, child = r.children(l)
        out[f"S{l}_parent_idx"], out[f"S{l}_off"], out[f"S{l}_child"] = pidx, off, child
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "sizes", [r.size(l) for l in range(levels + 1)])


if __name__ == "__main__":
    if not refhem.available():
        sys.exit("oracle/_ref/libhemref.so missing: run `make -C oracle ref` where /root/reference exists")
    for name in CASES:
        make(name)
