"""GPU parity of the ICP consumer against the CPU oracle (oracle/icp_oracle.py), through the
C ABI (hem_icp_point_to_point / hem_icp_multiscale).  Open3D itself is third-party and absent, so the oracle restates
its published RegistrationICP; tolerances: correspondences and iteration counts exact, fitness exact, inlier RMSE and
the transformation within 1e-9 (FP64 everywhere, only summation order differs)."""
import numpy as np
import pytest

from _icp import icp_oracle
from gaussiansplattingregistration_b200.synth import rigid_copy, synth

pytestmark = pytest.mark.gpu

TOL = 1e-9


def _rigid(rng, rot_deg, trans):
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    th = np.deg2rad(rot_deg)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = rng.normal(size=3) * trans
    return T


def _assert_same(g, o, what):
    assert g.iterations == o["iterations"], (what, g.iterations, o["iterations"])
    assert g.stop_reason == o["stop_reason"], (what, g.stop_reason, o["stop_reason"])
    assert g.n_correspondences == o["n_correspondences"], what
    assert g.fitness == o["fitness"], what
    assert abs(g.inlier_rmse - o["inlier_rmse"]) <= TOL * max(o["inlier_rmse"], 1e-12) + 1e-15, (what, g.inlier_rmse, o["inlier_rmse"])
    assert np.abs(g.transformation - o["transformation"]).max() <= TOL, (what, g.transformation, o["transformation"])
    if g.correspondence is not None:
        assert np.array_equal(g.correspondence, o["correspondence"]), (what, int((g.correspondence != o["correspondence"]).sum()))


@pytest.mark.parametrize("n,max_corr,noise,overlap", [(3000, 0.08, 0.0, 1.0), (20000, 0.05, 0.002, 1.0),
                                                       (20000, 0.03, 0.002, 0.6), (100000, 0.02, 0.001, 0.8)])
def test_icp_matches_oracle(hemctx, n, max_corr, noise, overlap):
    rng = np.random.default_rng(n + int(1000 * max_corr))
    tgt = rng.uniform(0, 1, (n, 3)) * np.array([2.0, 1.0, 0.5])
    T_true = _rigid(rng, 4.0, 0.01)
    keep = rng.uniform(size=n) < overlap
    src = tgt[keep]
    src = (src - T_true[:3, 3]) @ T_true[:3, :3] + rng.normal(size=src.shape) * noise     # T_true maps src onto tgt
    rng.shuffle(src)
    o = icp_oracle.registration_icp(src, tgt, max_corr, None, 1e-6, 1e-6, 40)
    g = hemctx.icp(src, tgt, None, max_corr, 1e-6, 1e-6, 40, with_correspondences=True)
    _assert_same(g, o, (n, max_corr))
    if noise == 0.0:
        assert np.abs(g.transformation - T_true).max() < 1e-8


def test_icp_init_and_zero_iterations(hemctx):
    rng = np.random.default_rng(5)
    tgt = rng.normal(size=(5000, 3))
    T0 = _rigid(rng, 2.0, 0.02)
    src = tgt + rng.normal(size=tgt.shape) * 0.01
    for iters in (0, 1, 7):
        o = icp_oracle.registration_icp(src, tgt, 0.3, T0, 1e-6, 1e-6, iters)
        g = hemctx.icp(src, tgt, T0, 0.3, 1e-6, 1e-6, iters, with_correspondences=True)
        _assert_same(g, o, iters)
    assert hemctx.icp(src, tgt, T0, 0.3, max_iteration=0).iterations == 0


def test_icp_edge_cases(hemctx):
    rng = np.random.default_rng(6)
    tgt = rng.uniform(0, 1, (2000, 3))
    # no overlap at all: no correspondences, transformation stays the initial one
    far = tgt + 100.0
    g = hemctx.icp(far, tgt, None, 0.05, max_iteration=10, with_correspondences=True)
    assert g.fitness == 0.0 and g.inlier_rmse == 0.0 and g.n_correspondences == 0 and g.stop_reason == 3
    assert np.array_equal(g.transformation, np.eye(4)) and (g.correspondence == -1).all()
    # empty clouds
    e = np.zeros((0, 3))
    for s, t in ((e, tgt), (tgt, e), (e, e)):
        g = hemctx.icp(s, t, None, 0.1, max_iteration=5, with_correspondences=True)
        assert g.fitness == 0.0 and g.n_correspondences == 0 and np.array_equal(g.transformation, np.eye(4))
    # planar and collinear targets (degenerate grid extents), duplicated points, a single target point
    planar = tgt * np.array([1.0, 1.0, 0.0])
    line = tgt * np.array([1.0, 0.0, 0.0])
    dup = np.repeat(tgt[:50], 40, axis=0)
    one = tgt[:1]
    for name, t in (("planar", planar), ("line", line), ("dup", dup), ("one", one)):
        s = t[: max(1, len(t) // 2)] + rng.normal(size=(max(1, len(t) // 2), 3)) * 1e-3
        o = icp_oracle.registration_icp(s, t, 0.05, None, 1e-6, 1e-6, 5)
        g = hemctx.icp(s, t, None, 0.05, max_iteration=5)
        assert g.n_correspondences == o["n_correspondences"] and g.fitness == o["fitness"], name
        assert abs(g.inlier_rmse - o["inlier_rmse"]) < 1e-6, name
    # non-finite points never match and never poison the sums
    s = tgt[:500].copy()
    s[::50] = np.nan
    t = tgt.copy()
    t[1999] = np.inf
    g = hemctx.icp(s, t, None, 0.05, max_iteration=3, with_correspondences=True)
    assert np.isfinite(g.transformation).all() and (g.correspondence[::50] == -1).all()
    assert g.n_correspondences == 500 - 10
    # huge search radius (every source point matches) and tiny radius
    o = icp_oracle.registration_icp(tgt[:300] + 0.01, tgt, 10.0, None, 1e-6, 1e-6, 3)
    g = hemctx.icp(tgt[:300] + 0.01, tgt, None, 10.0, max_iteration=3, with_correspondences=True)
    _assert_same(g, o, "huge radius")
    g = hemctx.icp(tgt[:300] + 0.01, tgt, None, 1e-9, max_iteration=3)
    assert g.n_correspondences == 0
    # argument errors come back as messages, not crashes
    from gaussiansplattingregistration_b200._capi import HemError
    with pytest.raises(HemError):
        hemctx.icp(tgt, tgt, None, 0.0)
    with pytest.raises(HemError):
        hemctx.icp(tgt, tgt, None, 0.1, max_iteration=-1)


def test_icp_is_deterministic_and_takes_device_pointers(hemctx):
    import torch
    rng = np.random.default_rng(8)
    tgt = rng.uniform(0, 1, (50000, 3))
    src = tgt[:30000] + rng.normal(size=(30000, 3)) * 0.003 + 0.004
    a = hemctx.icp(src, tgt, None, 0.03, max_iteration=20, with_correspondences=True)
    b = hemctx.icp(src, tgt, None, 0.03, max_iteration=20, with_correspondences=True)
    assert np.array_equal(a.transformation, b.transformation) and a.inlier_rmse == b.inlier_rmse
    assert np.array_equal(a.correspondence, b.correspondence)
    ds, dt = torch.from_numpy(src).cuda(), torch.from_numpy(tgt).cuda()
    h2d = hemctx.info("bytes_h2d")
    c = hemctx.icp(ds, dt, None, 0.03, max_iteration=20)
    assert hemctx.info("bytes_h2d") == h2d                                   # used in place
    assert np.array_equal(a.transformation, c.transformation) and a.fitness == c.fitness


def test_multiscale_icp_over_two_pyramids(port):
    """hem_icp_multiscale (two contexts, nothing leaves the GPU) == the oracle's coarse-to-fine loop on the same levels."""
    from gaussiansplattingregistration_b200.hem import HemContext
    from gaussiansplattingregistration_b200 import registration as reg
    a = synth(60000, seed=201, density=5000.0)
    b, T_true = rigid_copy(a, seed=202, max_rot_deg=6.0, max_trans_frac=0.02)
    L = 3
    spacing = (1.0 / 5000.0) ** (1.0 / 3.0)
    corr = [15 * spacing, 7.5 * spacing, 6 * spacing, 3 * spacing]           # coarse -> fine
    iters = [50, 30, 20, 10]                                                 # registration_parameters.py:8-15
    with HemContext(0, seed=13) as ca, HemContext(0, seed=13) as cb:
        ca.build(*a.arrays(), cluster_level=L)
        cb.build(*b.arrays(), cluster_level=L)
        pa = [ca.level(l)["xyz"].astype(np.float64) for l in range(L + 1)]
        pb = [cb.level(l)["xyz"].astype(np.float64) for l in range(L + 1)]
        got = reg.multiscale_registration(ca, cb, None, corr, iters)
        ref = icp_oracle.multiscale(pa, pb, corr, iters)
        assert len(got) == len(ref) == L + 1
        for i, (g, o) in enumerate(zip(got, ref)):
            _assert_same(g, o, ("scale", i))
        T = got[-1].transformation
        assert np.abs(T[:3, :3] - T_true[:3, :3]).max() < 2e-2 and np.abs(T[:3, 3] - T_true[:3, 3]).max() < 5e-2
        # the list form of the same driver (host arrays, one H2D per level) gives the same answer
        got2 = reg.multiscale_registration(pa, pb, None, corr, iters)
        for g, h in zip(got, got2):
            assert np.array_equal(g.transformation, h.transformation) and g.fitness == h.fitness
        # the reference's validity checks, same messages (qt_multiscale_registrator.py:170-194)
        from gaussiansplattingregistration_b200._capi import HemError
        with pytest.raises(HemError, match="mixture levels do not match"):
            ca.icp_multiscale(cb, corr[:2], iters[:2])
        cb.build(*b.arrays(), cluster_level=L - 1)
        with pytest.raises(HemError, match="differ in size"):
            ca.icp_multiscale(cb, corr, iters)


def test_do_icp_registration_mirror(hemctx):
    from gaussiansplattingregistration_b200 import registration as reg
    rng = np.random.default_rng(9)
    tgt = rng.uniform(0, 1, (8000, 3))
    src = tgt + 0.005

    class PC:                      # what an Open3D point cloud looks like to the caller
        def __init__(self, p):
            self.points = p
    p = reg.LocalRegistrationParams(max_correspondence=0.05, max_iteration=25)
    r = reg.do_icp_registration(PC(src), PC(tgt), np.eye(4), p, hemctx)
    o = icp_oracle.registration_icp(src, tgt, 0.05, np.eye(4), p.relative_fitness, p.relative_rmse, 25)
    _assert_same(r, o, "mirror")
    cs = r.correspondence_set
    assert cs.shape == (r.n_correspondences, 2) and np.array_equal(cs[:, 1], o["correspondence"][cs[:, 0]])
    with pytest.raises(NotImplementedError):
        reg.do_icp_registration(src, tgt, np.eye(4), reg.LocalRegistrationParams(registration_type=reg.LocalRegistrationType.ICP_Color), hemctx)


def test_register_gaussian_clouds_pipeline_matches_stepwise():
    """The one-call pipeline (two pyramids built concurrently + coarse-to-fine ICP) gives exactly what the steps give
    one after the other, and recovers the rigid motion."""
    from gaussiansplattingregistration_b200.hem import HemContext, HemParams
    from gaussiansplattingregistration_b200 import registration as reg
    a = synth(50000, seed=301, density=5000.0)
    b, T_true = rigid_copy(a, seed=302, max_rot_deg=5.0, max_trans_frac=0.02)
    spacing = (1.0 / 5000.0) ** (1.0 / 3.0)
    corr = [15 * spacing, 7.5 * spacing, 6 * spacing, 3 * spacing]
    iters = [50, 30, 20, 10]
    hp = HemParams(cluster_level=3)
    res, ca, cb = reg.register_gaussian_clouds(a.arrays(), b.arrays(), hp, corr, iters, seed=13)
    try:
        with HemContext(0, seed=13) as sa, HemContext(0, seed=13) as sb:
            sa.build(*a.arrays(), cluster_level=3)
            sb.build(*b.arrays(), cluster_level=3)
            for l in range(4):
                assert np.array_equal(sa.level(l)["xyz"], ca.level(l)["xyz"]) and np.array_equal(sb.level(l)["sh"], cb.level(l)["sh"])
            want = sa.icp_multiscale(sb, corr, iters)
        for g, w in zip(res, want):
            assert np.array_equal(g.transformation, w.transformation) and g.fitness == w.fitness and g.iterations == w.iterations
    finally:
        ca.close()
        cb.close()
    T = res[-1].transformation
    assert np.abs(T[:3, :3] - T_true[:3, :3]).max() < 2e-2 and np.abs(T[:3, 3] - T_true[:3, 3]).max() < 5e-2
    with pytest.raises(ValueError, match="mixture levels do not match"):
        reg.register_gaussian_clouds(a.arrays(), b.arrays(), hp, corr[:3], iters[:3])
