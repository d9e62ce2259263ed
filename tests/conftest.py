import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["hull4", "table1k", "ragged37", "single1", "edge2k"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 GPUs on the box (skipped otherwise)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def golden_path(name, kind):
    return os.path.join(GOLDEN, "%s.%s" % (name, kind))
