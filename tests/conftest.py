import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir("/root/reference/sparc_public")
    skip_ref = pytest.mark.skip(reason="/root/reference not present on this box")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def golden_ldpc():
    return np.load(os.path.join(GOLDEN, "ldpc_reference.npz"))


ALL_CODES = (
    [("802.16", r, z, p) for z in (3, 27, 54, 81)
     for r, p in (("1/2", "A"), ("2/3", "A"), ("2/3", "B"), ("3/4", "A"), ("3/4", "B"), ("5/6", "A"))]
    + [("802.11n", r, z, "A") for z in (27, 54, 81) for r in ("1/2", "2/3", "3/4", "5/6")]
)   # the 36 combinations of the reference's test_ldpc.py:5-42
