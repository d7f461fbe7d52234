import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLD = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gold():
    import numpy as np

    def load(name):
        p = GOLD / name
        if not p.exists():
            pytest.skip(f"golden fixture {name} not generated (tools/make_golden.py --full)")
        return np.load(p, allow_pickle=False)
    return load
