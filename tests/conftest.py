import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

def _ensure_built():
    """A fresh checkout has no built artefacts (they are git-ignored): build the product library, the drivers and the
    checker once, exactly as __graft_entry__.build() does.  Nothing is built when the files are already there."""
    import subprocess
    need = [os.path.join(ROOT, "p3arsec_b200", "lib", "libbs_gpu.so"), os.path.join(ROOT, "p3arsec_b200", "bin", "blackscholes_gpu"),
            os.path.join(ROOT, "p3arsec_b200", "bin", "bs_inputgen"), os.path.join(ROOT, "oracle", "libbs_oracle.so"),
            os.path.join(ROOT, "p3arsec_b200", "lib", "libsw_gpu.so"), os.path.join(ROOT, "p3arsec_b200", "bin", "swaptions_gpu"),
            os.path.join(ROOT, "oracle", "libsw_oracle.so")]
    if all(os.path.exists(p) for p in need):
        return
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "p3arsec_b200", "csrc"), "-j4", "all"], check=True)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "-j4", "all"], check=True)


_ensure_built()

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
