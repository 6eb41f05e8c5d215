import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


def rel_l2(a, b):
    import numpy as np
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def pytest_sessionfinish(session, exitstatus):
    """When the suite ran against a -DWT_BOUNDS_CHECK=1 build of the library (WT_LIB=...), fail the session if any
    shared-memory gather index of K1/K2 left its staged window (the stand-in for compute-sanitizer memcheck)."""
    if not os.environ.get("WT_LIB"):
        return
    try:
        import torch
        if not torch.cuda.is_available():
            return
        from whitomo_b200 import _capi
        n = int(_capi.load().wt_debug_oob_count())
    except Exception as e:          # the check is diagnostic; never mask the suite's own result
        print("\nwt_debug_oob_count unavailable: %r" % (e,))
        return
    print("\nwt_debug_oob_count = %d (library %s)" % (n, os.environ["WT_LIB"]))
    if n != 0:
        session.exitstatus = 1
