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
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    for k in list(d):
        if k.startswith("meta_") and d[k].dtype.kind in "US":
            d[k] = str(d[k])
    return d


def golden_names(prefix):
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith(prefix) and f.endswith(".npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import nlspn_oracle
    nlspn_oracle.build()
    return nlspn_oracle


@pytest.fixture
def nlspn_opt():
    """Sets tuning options of libnlspn_b200.so for one test (nlspn_set_option) and restores them afterwards."""
    from nlspn_eccv20_b200 import _lib
    old = {}

    def set_(**kw):
        for k, v in kw.items():
            old.setdefault(k, _lib.get_option(k))
            _lib.set_option(k, int(v))
    yield set_
    for k, v in old.items():
        _lib.set_option(k, v)
