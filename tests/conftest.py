import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure only)."""
    from oracle import oracle_py
    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def init_state(oracle):
    return oracle.read_bdimb(ROOT / "rlfluidcontrol_b200" / "data" / "init_state.bdimb")


@pytest.fixture(scope="session")
def rlfc():
    """The product package with librlfc.so built in-tree."""
    from rlfluidcontrol_b200 import build as _b
    _b.build()
    import rlfluidcontrol_b200
    return rlfluidcontrol_b200


def config1_actions(k):
    """BASELINE config 1 action sequence: a_k = (0.8 sin(2 pi k/25), -0.8 sin(2 pi k/25 + 1))."""
    return np.array([0.8 * np.sin(2 * np.pi * k / 25.0), -0.8 * np.sin(2 * np.pi * k / 25.0 + 1.0)], dtype=np.float32)
