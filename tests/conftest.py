import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_lib():
    import reflib
    reflib.build_port()
    return reflib.ora()


def data_folder(name):
    """tests/_data/<name>: built by the reference (tests/make_test_data.py). Built on demand where
    oracle/_ref exists; otherwise the test is skipped."""
    import reflib
    folder = os.path.join(reflib.DATA_DIR, name)
    if os.path.exists(os.path.join(folder, "indexloader.ini")) and os.path.exists(os.path.join(folder, "queries.npy")):
        return folder
    if not reflib.have_ref():
        pytest.skip("tests/_data/%s missing and oracle/_ref not built" % name)
    import make_test_data
    return make_test_data.make(name)
