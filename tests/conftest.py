import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def evr():
    import evr_sg4_b200
    evr_sg4_b200.lib.lib()      # builds libevr_sg4.so if needed and loads it
    return evr_sg4_b200


@pytest.fixture(scope="session")
def golden():
    import json
    g = {}
    d = os.path.join(ROOT, "tests", "golden")
    for name in ("sg4_tables", "kat", "herm_quadra"):
        with open(os.path.join(d, name + ".json")) as f:
            g[name] = json.load(f)
    return g
