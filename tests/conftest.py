import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# tests/test_emulated_kernels.py re-runs GPU parity tests in a subprocess over a SIMT-emulated g++ build of the kernel sources
# (tests/emu).  The hook that points the ctypes binding at that build lives HERE, in the test tree: the package has none.
if os.environ.get("RTB_TEST_EMULATION") == "1" and os.environ.get("RTB_EMU_LIB"):
    from raytracergpu_mastersproject_b200 import capi as _capi
    _capi._SO = os.environ["RTB_EMU_LIB"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def device():
    from raytracergpu_mastersproject_b200 import Device
    d = Device(0)
    yield d
    d.close()


@pytest.fixture
def make_device(monkeypatch):
    """Device factory for tests that set tuning knobs: librtb200 reads its RTB_WAVE_* environment knobs once, at rtb_ctx_create,
    so such a test sets the environment first and renders on its own context."""
    from raytracergpu_mastersproject_b200 import Device
    made = []

    def make(**env):
        for k, v in env.items():
            monkeypatch.setenv(k, str(v))
        d = Device(0)
        made.append(d)
        return d
    yield make
    for d in made:
        d.close()
