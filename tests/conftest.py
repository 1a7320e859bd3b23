import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIG_DIR = os.path.join(ROOT, "configs")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU check")


def config_path(name: str) -> str:
    return os.path.join(CONFIG_DIR, name if name.endswith(".yaml") else name + ".yaml")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def emu():
    from tests.emu import emu as E
    E.build()
    return E


@pytest.fixture(scope="session")
def gpu():
    """The product package, initialised on cuda:0.  GPU tests fail (not skip) when the extension is missing."""
    import r4w_b200
    r4w_b200.init(0)
    return r4w_b200
