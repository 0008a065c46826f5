import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU parity oracle (test infrastructure; built on demand with g++)."""
    from oracle import oracle_py

    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def shc_lib():
    """libshc_b200.so, built in-tree with nvcc if stale (cross-compiles without a GPU)."""
    from syropod_highlevel_controller_b200 import build, engine

    build.build()
    return engine.lib()
