"""Comparison helpers shared by the oracle tests and the GPU parity tests.

Parity rules: bit-exact for parent masks, per-parent child SETS,
orphan sets, counts and output order; FP32 tolerance 1e-4 relative for means,
covariances, weights, colours, opacities and SH, evaluated per attribute group
as ||a-b|| / max(||b||, tiny) because near-zero off-diagonals make element-wise
relative error meaningless.
"""
import numpy as np

TOL = 1e-4          # BASELINE.json north_star: "within 1e-4 relative in FP32"
ATTRS = ("xyz", "colors", "opacity", "cov6", "sh", "weight")


def group_relerr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin), "finite/non-finite pattern differs"
    if not fin.any():
        return 0.0
    return float(np.linalg.norm(a[fin] - b[fin]) / max(np.linalg.norm(b[fin]), 1e-30))


def assert_level_close(got, want, tol=TOL, what=""):
    assert got["xyz"].shape[0] == want["xyz"].shape[0], f"{what}: size {got['xyz'].shape[0]} != {want['xyz'].shape[0]}"
    for k in ATTRS:
        if k in got and k in want:
            e = group_relerr(got[k], want[k])
            assert e < tol, f"{what}: {k} group-relative error {e:.3e} >= {tol}"
    if "is_parent" in got and "is_parent" in want:
        assert np.array_equal(got["is_parent"], want["is_parent"]), f"{what}: parent flags differ"


def sort_children(off, child, *extra):
    """children ascending inside each parent; extras permuted alongside."""
    child = child.copy()
    extra = [e.copy() for e in extra]
    for s in range(len(off) - 1):
        a, b = int(off[s]), int(off[s + 1])
        if b - a > 1:
            o = np.argsort(child[a:b], kind="stable")
            child[a:b] = child[a:b][o]
            for e in extra:
                e[a:b] = e[a:b][o]
    return (child, *extra)


def assert_children_equal(got_pidx, got_off, got_child, want_pidx, want_off, want_child, what=""):
    assert np.array_equal(got_pidx, want_pidx), f"{what}: parent index lists differ"
    assert np.array_equal(np.asarray(got_off, np.int64), np.asarray(want_off, np.int64)), f"{what}: per-parent child counts differ"
    (gc,) = sort_children(got_off, np.asarray(got_child))
    (wc,) = sort_children(want_off, np.asarray(want_child))
    assert np.array_equal(gc, wc), f"{what}: child sets differ"


def inputs_of(level):
    return level["xyz"], level["colors"], level["opacity"], level["cov6"], level["sh"]
