import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # make sure the checker and the product library exist (compiling is not using)
    from oracle import pyoracle
    pyoracle.build()
    import montecarlolocalisation_b200 as m
    if not os.path.exists(m._lib.LIB_PATH):
        m.build()


@pytest.fixture(scope="session")
def map_txt():
    with open(os.path.join(ROOT, "tests", "golden", "map.txt")) as f:
        return f.read()
