import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_cases():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz"))


def slope_cases():
    """3-D fixtures with slope_type=1 (minmod) and slope_type=0 (no hydro slopes), tests/golden_slope/make_golden_slope.py."""
    d = os.path.join(ROOT, "tests", "golden_slope")
    return sorted(f[:-4] for f in os.listdir(d) if f.endswith(".npz"))


def init_only_cases():
    """Fixtures holding only the reference's step-0 state (problems the reference itself cannot step in 3-D: its rotor
    run produces NaN from the first step on, tests/golden/make_golden.py)."""
    d = os.path.join(GOLDEN, "init_only")
    return sorted("init_only/" + f[:-4] for f in os.listdir(d) if f.endswith(".npz")) if os.path.isdir(d) else []


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as O

    O.build()
    return O
