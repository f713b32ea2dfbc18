import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden", "ref")
GOLD_IN = os.path.join(GOLD, "in")
GOLD_OK = os.path.join(GOLD, "correct")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gold_in():
    return lambda name: os.path.join(GOLD_IN, name)


@pytest.fixture(scope="session")
def gold_ok():
    def _read(name):
        with open(os.path.join(GOLD_OK, name), "rb") as fh:
            return fh.read()
    return _read
