import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with `-m gpu` on a B200)")
    config.addinivalue_line("markers", "reference: needs the reference checkout (/root/reference)")


def pytest_collection_modifyitems(config, items):
    from tests.refharness import have_reference
    skip_ref = pytest.mark.skip(reason="reference checkout not present")
    for item in items:
        if "reference" in item.keywords and not have_reference():
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle
    oracle.build()
    return oracle.lib()
