import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ddp-generator_b200"))
sys.path.insert(0, os.path.join(ROOT, "ddp-generator_b200", "gen"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_devices():
    """Number of CUDA devices as the product library sees them (0 when no library is built or no GPU is present)."""
    try:
        import ilqg_b200

        return ilqg_b200.Library("car", 0).device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are the parity tests proper and need a B200; on a box without a CUDA device (CI, this container) they
    are skipped instead of failing, so a plain `pytest tests` stays green there.  On a GPU box nothing is skipped and a
    missing library still fails loudly."""
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (or product library not built): gpu-marked tests run on the B200 box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
