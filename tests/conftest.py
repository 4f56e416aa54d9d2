import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are selected with `-m gpu`; if someone runs them on a box without a CUDA
    device they must FAIL loudly (no silent skip, no CPU fallback)."""
    return


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
