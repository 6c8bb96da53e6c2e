"""Test session setup: register the ``gpu`` marker, make sure the native pieces are built
(idempotent, timestamp-checked) and expose small helpers shared by the parity tests."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    import __graft_entry__

    __graft_entry__.build()


@pytest.fixture(scope="session")
def have_ref():
    import oracle

    return oracle.have_ref()
