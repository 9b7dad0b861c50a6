import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # a gpu test on a machine without a GPU is a collection error we want to see
    # under -m gpu, not silently skip; under the default run they are deselected
    # by the driver with -m "not gpu".
    pass


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
