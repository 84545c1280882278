import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "nngp-src_b200", ROOT / "oracle"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"


@pytest.fixture(scope="session")
def mp_golden(golden_dir):
    import numpy as np
    z = np.load(golden_dir / "mp_small.npz")
    cases = {}
    for key in z.files:
        name, field = key.split("/")
        cases.setdefault(name, {})[field] = z[key]
    return cases


@pytest.fixture(scope="session")
def forest(golden_dir):
    import numpy as np
    z = np.load(golden_dir / "forest_xy.npz")
    return {k: z[k] for k in z.files}
