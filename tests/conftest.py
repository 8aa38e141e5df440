import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ast-text-analysis_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
ARRAY_NAMES = ["suftab", "lcptab", "childtab_up", "childtab_down", "childtab_next_l_index", "anntab"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json"), encoding="utf-8") as f:
        g = json.load(f)
    g["arrays"] = dict(np.load(os.path.join(GOLDEN_DIR, "golden_arrays.npz")))
    return g


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.build()
    return oracle


def have_gpu():
    try:
        from east import _capi
        return _capi.device_count() > 0
    except Exception:
        return False


@pytest.fixture(autouse=True)
def _no_alphabet_guess_across_tests(request):
    """A batch of small documents starts from the alphabet of the thread's previous batch (a guess, see
    test_alphabet_of_the_previous_batch_is_only_a_guess).  Tests that look at the miss statistics of an index must
    not depend on which test ran before them: every GPU test starts without a guess."""
    if request.node.get_closest_marker("gpu") is not None:
        from east import _capi
        _capi.set_option("forget_alphabet_guess", 1)
    yield
