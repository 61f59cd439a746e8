import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


def params_of(g, prefix="p."):
    return {k[len(prefix):]: v.copy() for k, v in g.items() if k.startswith(prefix)}


def rel_err(a, b):
    """max |a-b| / max(|b|, tiny): normwise relative error used by all parity tests."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0 and b.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def elem_err(a, b, floor=1e-3):
    """Elementwise relative error with an absolute floor: max |a-b| / (|b| + floor * max|b|).  Tighter than `rel_err`
    wherever one large component (the residual observation) would hide an error in a small one."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0 and b.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / (np.abs(b) + floor * max(float(np.max(np.abs(b))), 1e-30))))


_MAXIMA = {}


def record_max(name, value):
    """Measured parity maxima, written at session end to gpurun_out/parity_maxima.txt (copied to profiles/ per round)."""
    _MAXIMA[name] = max(_MAXIMA.get(name, 0.0), float(value))
    return value


def pytest_sessionfinish(session, exitstatus):
    if not _MAXIMA:
        return
    out_dir = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "parity_maxima.txt"), "w") as f:
            f.write("# measured maxima of the parity tests (normwise rel_err unless the name says elem)\n")
            for k in sorted(_MAXIMA):
                f.write("%-90s %.3e\n" % (k, _MAXIMA[k]))
    except OSError:
        pass
