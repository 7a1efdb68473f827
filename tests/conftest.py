"""pytest wiring: markers, import paths for the fakes / oracle, native-library build."""
from __future__ import annotations

import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "tests", ROOT / "tests" / "fakes", ROOT / "oracle"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

# The CPU suite drives the simulated register file; a GPU box picks cudasim itself.
os.environ.setdefault("CC_READINESS_FILE", "/tmp/ccm-test-readiness/.cc-manager-ctr-ready")
# The manager resets a GPU's CUDA primary context after the scrub gate (production default).
# Inside ONE pytest process that context is shared with torch and with later tests, so the suite
# keeps contexts alive; the release path is exercised in a subprocess (tests/test_context_release.py).
os.environ.setdefault("CC_RELEASE_CUDA_CONTEXT", "false")
# manager.main() refuses libccm's simulated register backends unless told the drill is intended
os.environ.setdefault("CCM_ALLOW_SIM", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_library():
    """libccm.so must exist before anything imports the product; build it if the
    tree is fresh (nvcc cross-compiles without a GPU)."""
    from k8s_cc_manager_b200 import build
    build.build()
    yield


@pytest.fixture()
def cluster():
    import kubernetes
    return kubernetes.reset_cluster()


@pytest.fixture()
def native():
    from k8s_cc_manager_b200 import _native
    return _native


def cuda_device_count() -> int:
    import ctypes as C
    from k8s_cc_manager_b200 import _native as N
    lib = N.lib()
    lib.ccm_init(N.BACKEND_CUDASIM)
    n = C.c_int(0)
    lib.ccm_enumerate(None, 0, C.byref(n))
    return n.value
