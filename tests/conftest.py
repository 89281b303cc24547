import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def _cuda_devices():
    """CUDA devices the product library sees (0 when the library is not built or no GPU is present)."""
    try:
        import xsparse_b200

        return xsparse_b200.capi.device_count()
    except Exception:  # noqa: BLE001
        return 0


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without a GPU skips the gpu-marked tests instead of failing them in
    xsb_create ("no CUDA device: libxsparse_b200 has no CPU fallback")."""
    if not any("gpu" in item.keywords for item in items):
        return
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (libxsparse_b200 has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as ora

    ora.build()
    return ora
