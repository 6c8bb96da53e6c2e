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
    import oracle
    from mc_dagprop_b200 import build

    build.build_all()  # timestamp-checked; __graft_entry__.build() is the from-source check
    oracle.build()


@pytest.fixture(scope="session")
def have_ref():
    import oracle

    return oracle.have_ref()
