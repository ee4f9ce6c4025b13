import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sbv2-api_b200"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def lib_built():
    """Builds libsbv2_b200.so if it is not there (CPU box: nvcc cross-compiles)."""
    so = os.path.join(ROOT, "sbv2-api_b200", "sbv2_b200", "libsbv2_b200.so")
    if not os.path.exists(so):
        import importlib.util
        spec = importlib.util.spec_from_file_location("sbv2_b200_build", os.path.join(ROOT, "sbv2-api_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    return so


def _has_gpu():
    try:
        import sbv2_b200 as S
        return S.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # GPU tests fail loudly (not skip) when selected with -m gpu on a box without a device; when
    # the whole suite is run unfiltered on a CPU box they are skipped.
    selected = config.getoption("-m") or ""
    if "gpu" in selected and "not gpu" not in selected:
        return
    if not _has_gpu():
        skip = pytest.mark.skip(reason="no B200 visible")
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)
