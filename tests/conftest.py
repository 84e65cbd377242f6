import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


# fixtures that hold summaries (crops, means) of outputs too large to store; they have their own tests
SUMMARY_FIXTURES = ("cfg3_2160x3840_gauss63_n200", "fanout_", "restorer_", "act_u8")


def golden_names(prefix=None, exclude=()):
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        n = os.path.basename(f)[:-4]
        if n.startswith(SUMMARY_FIXTURES) and not (prefix and n.startswith(prefix)):
            continue
        if prefix and not n.startswith(prefix):
            continue
        if any(n.startswith(e) for e in exclude):
            continue
        out.append(n)
    return out


@pytest.fixture(scope="session")
def lib():
    """Built C-ABI library (builds it if the sources are newer)."""
    from torch_admm_deconv_b200 import build, _lib
    build.build()
    return _lib.load()
