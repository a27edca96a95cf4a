import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from _oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from _oracle import Reference, have_reference
    if not have_reference():
        pytest.skip("oracle/_ref/libref.so not built (reference tree absent)")
    return Reference()


@pytest.fixture(scope="session")
def revised():
    from _oracle import RevisedReference, have_revised_reference
    if not have_revised_reference():
        pytest.skip("oracle/_ref/libref_revised.so not built (reference tree absent)")
    return RevisedReference()


@pytest.fixture(scope="session")
def sp():
    import superman_b200
    return superman_b200
