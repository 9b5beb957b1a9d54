import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def port():
    from oracle.oracle import Port, build
    build(ref=True)
    return Port()


@pytest.fixture(scope="session")
def ref():
    from oracle.oracle import Reference, REF_SO
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return Reference()


@pytest.fixture(scope="session")
def reftest():
    from oracle.oracle import RefTestUtils, REFTEST_SO
    if not os.path.exists(REFTEST_SO):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return RefTestUtils()
