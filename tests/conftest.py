import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than ~20 s on CPU")


def _pf(v):
    if isinstance(v, str):
        return {"nan": float("nan"), "inf": float("inf"), "-inf": float("-inf")}[v]
    return float(v)


def parse_floats(a):
    return np.array([_pf(v) for v in a], dtype=np.float64)


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_golden.json")) as fh:
        g = json.load(fh)
    g["_pf"] = _pf
    return g


@pytest.fixture(scope="session")
def golden2():
    """Round-2 reference goldens (oracle/make_golden.py --r2)."""
    with open(os.path.join(ROOT, "tests", "golden", "reference_golden_r2.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def golden_ppc_onebd():
    """utilities/ppcTools_oneBD.py goldens (oracle/make_golden.py --ppc-onebd)."""
    with open(os.path.join(ROOT, "tests", "golden", "reference_golden_ppc_onebd.json")) as fh:
        return json.load(fh)["ppc_onebd"]


def unsparse(d):
    a = np.zeros(int(np.prod(d["shape"])), dtype=np.int64)
    a[np.asarray(d["idx"], dtype=np.int64)] = np.asarray(d["val"], dtype=np.int64)
    return a.reshape(d["shape"])


@pytest.fixture(scope="session")
def pf():
    return _pf
