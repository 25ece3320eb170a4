"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every
symbol include/hem_b200.h declares, the Python binding covers them, the product
path fails loudly without a GPU, and the mixture_bind shim's host logic mirrors
the reference's argument handling.  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "hem_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hem_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from gaussiansplattingregistration_b200 import _capi
    names = declared_functions()
    assert len(names) >= 12
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hem_b200.h but not exported"
    assert set(names) == set(_capi.SIGNATURES), "python binding and header disagree"
    assert _capi.load().hem_version() >= 1000


def test_product_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gaussiansplattingregistration_b200 import _capi
    from gaussiansplattingregistration_b200.hem import HemContext
    with pytest.raises(_capi.HemError):
        HemContext(0)


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "gaussiansplattingregistration_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports oracle"
                assert "hem_oracle" not in txt and "libhemref" not in txt, f"{f} references the oracle"


def test_mixture_bind_surface_and_marshalling():
    from gaussiansplattingregistration_b200 import mixture_bind as mb
    # names and constructors of reference mixture_bind.cpp:14-56
    v = mb.vec3([1, 2, 3])
    assert (v.x, v.y, v.z) == (1.0, 2.0, 3.0) and repr(mb.vec3()).startswith("<vec3(")
    m = mb.smat3(1, 2, 3, 4, 5, 6)
    assert m.e12 == 5.0 and mb.smat3([1, 2, 3, 4, 5, 6]).e22 == 6.0
    with pytest.raises(RuntimeError, match="exactly 3 elements"):
        mb.vec3([1, 2])
    with pytest.raises(RuntimeError, match="exactly 6 elements"):
        mb.smat3([1, 2, 3])
    assert mb.FeatureVector(4).GetSize() == 4 and mb.FeatureVector([1.0, 2.0]).GetVector() == [1.0, 2.0]
    # CreateMixtureLevel argument order: xyz, colors, opacities, covariance, features (mixturelevel.cpp:14-28)
    n, F = 5, 9
    xyz = np.arange(n * 3, dtype=np.float32).reshape(n, 3)
    col = xyz + 100
    op = np.arange(n, dtype=np.float32)
    cov = np.arange(n * 6, dtype=np.float32).reshape(n, 6)
    sh = np.arange(n * F, dtype=np.float32).reshape(n, F)
    lvl = mb.MixtureLevel.CreateMixtureLevel(xyz.tolist(), col.tolist(), op.tolist(), cov.tolist(), sh.tolist())
    assert len(lvl) == n and lvl.features.shape == (n, F)
    lists = mb.MixtureLevel.CreatePythonLists(lvl)
    assert isinstance(lists, tuple) and len(lists) == 5 and isinstance(lists[0][0], list)
    for got, want in zip(lists, (xyz, col, op, cov, sh)):
        assert np.array_equal(np.asarray(got, np.float32), want)
    # wrong inner lengths raise like the reference (vec.hpp:92-94, 473-475)
    with pytest.raises(RuntimeError, match="exactly 3 elements"):
        mb.MixtureLevel.CreateMixtureLevel([[0, 0]], [[0, 0, 0]], [0], [[0] * 6], [[0]])
    with pytest.raises(RuntimeError, match="exactly 6 elements"):
        mb.MixtureLevel.CreateMixtureLevel([[0, 0, 0]], [[0, 0, 0]], [0], [[0] * 5], [[0]])
    # clusterLevel 0 returns [] without touching the GPU (mixture_wrapper.cpp:14-15 erases level 0)
    assert mb.MixtureCreator.CreateMixture(0, 3.0, 3.0, 2.5, 1.0, lvl) == []


def test_synth_is_deterministic():
    from gaussiansplattingregistration_b200.synth import synth
    a, b = synth(1000, seed=5), synth(1000, seed=5)
    for x, y in zip(a.arrays(), b.arrays()):
        assert np.array_equal(x, y)
    assert a.sh.shape == (1000, 45)
    # covariances are positive definite
    c = a.cov6.astype(np.float64)
    det = (-c[:, 2] * c[:, 2] * c[:, 3] + 2 * c[:, 1] * c[:, 2] * c[:, 4] - c[:, 0] * c[:, 4] * c[:, 4]
           - c[:, 1] * c[:, 1] * c[:, 5] + c[:, 0] * c[:, 3] * c[:, 5])
    assert (det > 0).all()


def test_registration_host_side_types():
    """Host-only pieces of the ICP mirror: result unpacking, Open3D-style correspondence_set, point extraction,
    parameter defaults equal to the reference's (src/params/registration_parameters.py:8-15)."""
    import numpy as np
    from gaussiansplattingregistration_b200 import _capi
    from gaussiansplattingregistration_b200 import registration as reg
    from gaussiansplattingregistration_b200.hem import IcpResult
    flat = np.zeros(_capi.ICP_RESULT_LEN)
    flat[:16] = np.eye(4).ravel()
    flat[16:24] = [0.75, 0.01, 7, 3, 1, 0.5, 1000, 0.1]
    corr = np.array([2, -1, 0, 5], np.int32)
    r = IcpResult.from_flat(flat, corr)
    assert r.fitness == 0.75 and r.inlier_rmse == 0.01 and r.iterations == 7 and r.n_correspondences == 3
    assert r.stop_reason == 1 and r.grid_cells == 1000 and np.array_equal(r.transformation, np.eye(4))
    assert np.array_equal(r.correspondence_set, np.array([[0, 2], [2, 0], [3, 5]], np.int32))
    with pytest.raises(ValueError):
        IcpResult.from_flat(flat, None).correspondence_set

    class PC:
        points = [[0.0, 1.0, 2.0], [3.0, 4.0, 5.0]]
    for src in (PC(), {"points": PC.points}, PC.points, np.asarray(PC.points, np.float32)):
        p = reg._points_of(src)
        assert p.dtype == np.float64 and p.shape == (2, 3) and p.flags["C_CONTIGUOUS"]
    d = reg.LocalRegistrationParams()
    assert (d.max_correspondence, d.relative_fitness, d.relative_rmse, d.max_iteration, d.k_value) == (5.0, 1e-6, 1e-6, 30, 0.0)
    assert d.rejection_type is reg.KernelLossFunctionType.Loss_None
    assert [t.instance_name for t in reg.LocalRegistrationType] == ["Point-to-Point ICP", "Point-to-Plane ICP", "Colored ICP",
                                                                    "Generalized ICP"]
    with pytest.raises(NotImplementedError):      # raised before any GPU work
        reg.do_icp_registration(PC(), PC(), np.eye(4), reg.LocalRegistrationParams(registration_type=reg.LocalRegistrationType.ICP_General))
    with pytest.raises(ValueError, match="mixture levels do not match"):
        reg.register_gaussian_clouds((), (), None, [1.0], [5])


def test_options_and_statistics_are_documented():
    """Every option hem_set_option accepts is described in include/hem_b200.h, and the step statistics have the same
    length in the header, the ctypes binding and the Python names."""
    from gaussiansplattingregistration_b200 import _capi
    from gaussiansplattingregistration_b200.hem import STAT_NAMES
    header = open(HEADER).read()
    host = open(os.path.join(ROOT, "gaussiansplattingregistration_b200", "csrc", "hem_b200.cu")).read()
    body = host[host.index("int hem_set_option("):]
    body = body[:body.index("\n}\n")]
    options = set(re.findall(r'k == "([a-z0-9_]+)"', body))
    assert len(options) >= 15
    undocumented = sorted(o for o in options if f'"{o}"' not in header)
    assert not undocumented, f"options missing from include/hem_b200.h: {undocumented}"
    m = re.search(r"#define\s+HEM_STEP_STATS_LEN\s+(\d+)", header)
    assert m and int(m.group(1)) == _capi.STEP_STATS_LEN == len(STAT_NAMES)
