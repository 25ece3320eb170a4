"""CPU tests: the oracle restatement (oracle/hem_oracle.c) against the golden
fixtures recorded from the real reference, against the reference itself where
oracle/_ref exists, and against the invariants the reference's algorithm obeys
(SURVEY.md §4)."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden
from _parity import assert_children_equal, assert_level_close, inputs_of, sort_children
from gaussiansplattingregistration_b200.synth import synth

# Philox4x32-10 known-answer vectors (Random123 kat_vectors), first output word
PHILOX_KAT = [
    (0x0000000000000000, 0x0000000000000000, 0x6627E8D5),
    (0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF, None),      # counter words c2,c3 are 0 here, so only self-consistency
]


def test_philox_known_answer(port):
    assert port.philox_u32(0, 0) == 0x6627E8D5
    # determinism + sensitivity to both key and counter
    assert port.philox_u32(1, 0) != port.philox_u32(0, 0)
    assert port.philox_u32(0, 1) != port.philox_u32(0, 0)
    assert port.philox_u32(0x123456789ABCDEF0, 1 << 40) == port.philox_u32(0x123456789ABCDEF0, 1 << 40)


def _run_golden(port, g):
    red, dd, cd, decay = (float(v) for v in g["params"])
    return port.run(g["xyz"], g["colors"], g["opacity"], g["cov6"], g["sh"], int(g["levels"]), red, dd, cd, decay,
                    seed=int(g["seed"]))


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_reference_golden(port, name):
    """Free-running pyramid vs the recorded reference run: counts, order, flags and
    child sets exact; values within tolerance."""
    g = load_golden(name)
    r = _run_golden(port, g)
    L = int(g["levels"])
    for l in range(L + 1):
        want = {k: g[f"L{l}_{k}"] for k in ("xyz", "colors", "opacity", "cov6", "sh", "weight", "is_parent")}
        assert_level_close(r.level(l), want, what=f"{name} level {l}")
    for l in range(L):
        c = r.children(l)
        assert_children_equal(c["parent_idx"], c["off"], c["child"], g[f"S{l}_parent_idx"], g[f"S{l}_off"],
                              g[f"S{l}_child"], what=f"{name} step {l}")


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_teacher_forced_on_golden(port, name):
    """One step from the REFERENCE's own level-l rows, weights and flags must
    reproduce the reference's level l+1 exactly in membership (SURVEY.md §7 mode A)."""
    g = load_golden(name)
    red, dd, cd, decay = (float(v) for v in g["params"])
    L = int(g["levels"])
    base = int(g["n"])
    for l in range(L):
        lvl = {k: g[f"L{l}_{k}"] for k in ("xyz", "colors", "opacity", "cov6", "sh", "weight", "is_parent")}
        r = port.run(*inputs_of(lvl), 1, red, dd, cd, decay, seed=int(g["seed"]), weight=lvl["weight"],
                     mask=lvl["is_parent"], rng_base=base)
        c = r.children(0)
        assert_children_equal(c["parent_idx"], c["off"], c["child"], g[f"S{l}_parent_idx"], g[f"S{l}_off"],
                              g[f"S{l}_child"], what=f"{name} step {l}")
        want = {k: g[f"L{l+1}_{k}"] for k in ("xyz", "colors", "opacity", "cov6", "sh", "weight", "is_parent")}
        assert_level_close(r.level(1), want, what=f"{name} level {l+1} (teacher-forced)")
        base += r.stats(0)["nprefilter"]


def test_oracle_vs_reference_live(port):
    """Where the compiled reference is present (oracle/_ref), compare live on a fresh seed."""
    from oracle import refhem
    if not refhem.available():
        pytest.skip("oracle/_ref/libhemref.so not present")
    c = synth(6000, seed=21, density=5000)
    ref = refhem.run(*c.arrays(), levels=3, seed=5)
    r = port.run(*c.arrays(), levels=3, seed=5)
    for l in range(4):
        assert_level_close(r.level(l), ref.level(l), what=f"level {l}")
    for l in range(3):
        pidx, off, child = ref.children(l)
        pc = r.children(l)
        assert_children_equal(pc["parent_idx"], pc["off"], pc["child"], pidx, off, child, what=f"step {l}")
        if l == 0:      # identical inputs only at level 0: query radius (mixture.cpp:88, vec.hpp:736-768) bit for bit
            assert np.array_equal(pc["radius"], ref.radii(l)[pidx])


def test_oracle_invariants(port):
    """Mass conservation, weights >= 1, parents are their own children, orphans are
    bit-copies, output order = parents ascending then orphans ascending."""
    c = synth(8000, seed=3, density=5000)
    r = port.run(*c.arrays(), levels=3, seed=9)
    for l in range(3):
        a, b = r.level(l), r.level(l + 1)
        st = r.stats(l)
        ch = r.children(l)
        assert st["ndropped"] == 0
        assert b["xyz"].shape[0] == st["nparents"] + st["norphans"] < a["xyz"].shape[0]
        assert abs(float(b["weight"].astype(np.float64).sum()) - 8000.0) < 8000 * 1e-5
        assert float(b["weight"].min()) >= 1.0 - 1e-5
        for s, p in enumerate(ch["parent_idx"]):
            assert p in ch["child"][ch["off"][s]:ch["off"][s + 1]]
        orphans = np.nonzero(ch["sumlw"] == 0.0)[0]
        assert len(orphans) == st["norphans"]
        np_ = st["nparents"]
        for k in ("xyz", "colors", "opacity", "cov6", "sh", "weight"):
            assert np.array_equal(b[k][np_:], a[k][orphans])
        assert np.all(np.diff(ch["parent_idx"].astype(np.int64)) > 0)


def test_oracle_edge_cases(port):
    c = synth(500, seed=4, density=5000)
    # cluster level 0: only the input level
    r = port.run(*c.arrays(), levels=0, seed=1)
    assert r.num_levels == 1 and r.size(0) == 500
    # no parent selected: every component is an orphan, the level is a copy (in order)
    r = port.run(*c.arrays(), levels=2, seed=1, mask=np.zeros(500, np.uint8), rng_base=500)
    a, b = r.level(0), r.level(1)
    assert r.stats(0)["nparents"] == 0 and r.stats(0)["norphans"] == 500
    for k in ("xyz", "colors", "opacity", "cov6", "sh", "weight"):
        assert np.array_equal(a[k], b[k])
    # empty input
    e = synth(0, seed=0)
    r = port.run(*e.arrays(), levels=2, seed=1)
    assert [r.size(l) for l in range(3)] == [0, 0, 0]
    # every component a parent: each keeps only itself (parent exclusion, mixture.cpp:131-133)
    r = port.run(*c.arrays(), levels=1, seed=1, mask=np.ones(500, np.uint8), rng_base=500)
    ch = r.children(0)
    assert np.array_equal(ch["child"], np.arange(500, dtype=np.uint32))
    assert_level_close({k: v for k, v in r.level(1).items() if k != "is_parent"},
                       {k: v for k, v in r.level(0).items() if k != "is_parent"}, tol=1e-5)


def test_icp_oracle_recovers_a_known_rigid_motion_and_validates_arguments():
    """The ICP oracle (restated Open3D RegistrationICP, parity unpinned by the reference) on a noiseless copy, and the
    host-side validity checks of the registration mirror (no GPU involved)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
    import icp_oracle
    from gaussiansplattingregistration_b200 import registration as reg
    rng = np.random.default_rng(3)
    tgt = rng.uniform(0, 1, (4000, 3))
    th = 0.04
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    t = np.array([0.01, -0.015, 0.02])
    src = (tgt - t) @ R                     # R src + t = tgt
    r = icp_oracle.registration_icp(src, tgt, 0.1, None, 1e-6, 1e-6, 50)
    assert r["fitness"] == 1.0 and r["inlier_rmse"] < 1e-12 and r["stop_reason"] == 1
    assert np.abs(r["transformation"][:3, :3] - R).max() < 1e-12 and np.abs(r["transformation"][:3, 3] - t).max() < 1e-12
    assert np.array_equal(r["correspondence"], np.arange(4000))
    # a reflection-prone (planar) configuration still yields a proper rotation
    U = icp_oracle.umeyama_rigid(src * [1, 1, 0], tgt * [1, 1, 0])
    assert abs(np.linalg.det(U[:3, :3]) - 1.0) < 1e-12
    with pytest.raises(ValueError, match="do not match"):
        reg.multiscale_registration([src, src], [tgt, tgt], None, [0.1, 0.05], [10])
    with pytest.raises(ValueError, match="differ in size"):
        reg.multiscale_registration([src, src], [tgt], None, [0.1, 0.05], [10, 5])
    with pytest.raises(ValueError, match="no downscaled mixtures"):
        reg.multiscale_registration([src], [tgt], None, [0.1], [10])
    with pytest.raises(ValueError, match="mixture levels do not match"):
        reg.multiscale_registration([src, src, src], [tgt, tgt, tgt], None, [0.1, 0.05], [10, 5])
    assert reg.LocalRegistrationParams().registration_type is reg.LocalRegistrationType.ICP_Point_To_Point


def test_oracle_vs_reference_random_sweep(port):
    """Oracle pinning beyond the fixed fixtures: 24 random small clouds with random CreateMixture arguments,
    densities, SH degrees and injected non-finite / indefinite covariances, first step compared with the compiled
    reference (child sets, masks, counts, order exact; values to summation order)."""
    from oracle import refhem
    if not refhem.available():
        pytest.skip("oracle/_ref/libhemref.so not present")
    rng = np.random.default_rng(2026)
    for case in range(24):
        n = int(rng.integers(50, 1500))
        deg = int(rng.integers(0, 5))
        c = synth(n, seed=1000 + case, density=float(rng.choice([500.0, 5000.0, 50000.0])), sh_degree=deg)
        red = float(rng.choice([1.0, 1.5, 3.0, 8.0]))
        dd = float(rng.choice([0.5, 2.0, 3.0, 6.0]))
        cd = float(rng.choice([0.0, 1.0, 2.5, 10.0]))
        decay = float(rng.choice([0.05, 1.0, 30.0]))
        if case % 4 == 3:                     # poison a few components the way real exports do
            k = rng.integers(0, n, size=5)
            c.cov6[k[0]] = 0.0
            c.cov6[k[1], 0] = -abs(c.cov6[k[1], 0])
            c.cov6[k[2], 3] = np.nan
            c.xyz[k[3], 1] = np.nan
            c.opacity[k[4]] = np.inf
        seed = 77 + case
        ref = refhem.run(*c.arrays(), 1, red, dd, cd, decay, seed=seed)
        r = port.run(*c.arrays(), 1, red, dd, cd, decay, seed=seed)
        what = f"case {case}: n={n} deg={deg} red={red} dd={dd} cd={cd} decay={decay}"
        assert r.size(1) == ref.size(1), what
        pidx, off, child = ref.children(0)
        pc = r.children(0)
        assert_children_equal(pc["parent_idx"], pc["off"], pc["child"], pidx, off, child, what=what)
        a, b = r.level(1), ref.level(1)
        assert np.array_equal(a["is_parent"], b["is_parent"]), what
        for k in ("xyz", "colors", "opacity", "cov6", "sh", "weight"):
            fa, fb = np.isfinite(a[k]), np.isfinite(b[k])
            assert np.array_equal(fa, fb), (what, k)
            m = fa & fb
            num = np.linalg.norm((a[k][m].astype(np.float64) - b[k][m]))
            den = max(np.linalg.norm(b[k][m].astype(np.float64)), 1e-30)
            assert num / den < 1e-4, (what, k, num / den)
