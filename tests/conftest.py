import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # Build the checker (and, in the build container, oracle/_ref) once per session.
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)


def _has_gpu() -> bool:
    try:
        import lzbench_b200 as L
        return L.lib.lz4b200_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_ctx():
    import lzbench_b200 as L
    if not _has_gpu():
        pytest.fail("GPU test selected but no CUDA device / library: the lz4_b200 path has no CPU fallback")
    ctx = L.Lz4B200(0)
    yield ctx
    ctx.close()
