import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the in-tree artefacts exist (the driver runs build() before the tests; this
    covers a bare `pytest`).  Never rebuilds what is already there."""
    need = [
        os.path.join(ROOT, "k3s-nvidia_b200", "libb200probe.so"),
        os.path.join(ROOT, "oracle", "liboracle.so"),
        os.path.join(ROOT, "tests", "mock_nvml", "libnvidia-ml-mock.so"),
        os.path.join(ROOT, "tests", "fake_probe", "libfakeprobe.so"),
        os.path.join(ROOT, "host", "cpp", "build", "b200-device-plugin"),
    ]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__ as g

        g.build()
    yield
