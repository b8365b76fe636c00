import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle_lib():
    from tests import oracle
    return oracle.lib()


@pytest.fixture(scope="session")
def ctx():
    """One bt_context for the GPU tests; fails loudly (no skip, no fallback) if the C-ABI library or the device is missing."""
    from bevy_terrain_b200 import api
    c = api.Context(0)
    yield c
    c.close()
