import importlib
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "timeout: per-test limit in seconds (pytest-timeout; ignored if the plugin is absent)")


@pytest.fixture(scope="session")
def problems():
    return importlib.import_module("tinympc-matlab_b200.problems")


@pytest.fixture(scope="session")
def oracle_mod():
    """The CPU checker (test infrastructure). Builds the C port on demand."""
    port = ROOT / "oracle" / "liboracle_port.so"
    if not port.exists():
        subprocess.check_call(["make", "-C", str(ROOT / "oracle"), "liboracle_port.so"])
    import oracle  # noqa
    return oracle
