"""Downstream parity (BASELINE configs[4] in miniature): coarse-to-fine ICP over the pyramids of two
clouds, built once by the oracle and once by the CUDA path, must agree within 1e-3 in fitness and
inlier RMSE at every scale and in the recovered transformation."""
import numpy as np
import pytest

from _icp import multiscale_icp
from gaussiansplattingregistration_b200.synth import rigid_copy, synth

pytestmark = pytest.mark.gpu


def test_multiscale_icp_on_gpu_pyramids_matches_oracle_pyramids(port):
    from gaussiansplattingregistration_b200.hem import HemContext
    a = synth(40000, seed=101, density=5000.0)
    b, T_true = rigid_copy(a, seed=102, max_rot_deg=6.0, max_trans_frac=0.02)
    L = 3
    pyr = {}
    with HemContext(0, seed=13) as ctx:
        for name, c in (("a", a), ("b", b)):
            ctx.build(*c.arrays(), cluster_level=L)
            pyr["gpu_" + name] = [ctx.level(l)["xyz"] for l in range(L + 1)]
            r = port.run(*c.arrays(), levels=L, seed=13)
            pyr["cpu_" + name] = [r.level(l)["xyz"] for l in range(L + 1)]
    spacing = (1.0 / 5000.0) ** (1.0 / 3.0)
    corr = [5 * spacing * 3, 2.5 * spacing * 3, 2 * spacing * 3, 1 * spacing * 3]     # coarse -> fine
    iters = [50, 30, 20, 10]                                                         # registration_parameters.py:8-15
    Tg, sg = multiscale_icp(pyr["gpu_a"], pyr["gpu_b"], corr, iters)
    Tc, sc = multiscale_icp(pyr["cpu_a"], pyr["cpu_b"], corr, iters)
    for (fg, rg), (fc, rc) in zip(sg, sc):
        assert abs(fg - fc) <= 1e-3, (sg, sc)
        assert abs(rg - rc) <= 1e-3 * max(rc, 1e-12) + 1e-9, (sg, sc)
    assert np.abs(Tg - Tc).max() < 1e-3
    # and the registration actually works: the known rigid motion is recovered at the finest scale
    assert np.abs(Tg[:3, :3] - T_true[:3, :3]).max() < 2e-2 and np.abs(Tg[:3, 3] - T_true[:3, 3]).max() < 5e-2, (Tg, T_true)
