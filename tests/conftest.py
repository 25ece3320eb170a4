import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["dense_sh3", "sparse_sh1_params", "degenerate_sh0"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def port():
    """The plain-C oracle (oracle/hem_oracle.c), built on demand."""
    from oracle import porthem
    porthem.lib()
    return porthem


@pytest.fixture(scope="session")
def hemctx():
    """One CUDA context for the GPU tests; fails loudly if the extension or GPU is missing."""
    from gaussiansplattingregistration_b200.hem import HemContext
    ctx = HemContext(0, seed=7)
    ctx.set_option("keep_pairs", 1)
    yield ctx
    ctx.close()
