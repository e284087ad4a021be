import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` under gpurun)")


@pytest.fixture(scope="session")
def built_lib():
    """liboar_b200.so built in-tree (nvcc cross-compiles without a GPU)."""
    from oar_ocr_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def ctx(built_lib):
    from oar_ocr_b200 import ffi
    return ffi.Context(0)


@pytest.fixture(scope="session")
def det_blob():
    from oar_ocr_b200 import models
    return models.get_blob("det")


@pytest.fixture(scope="session")
def rec_blob():
    from oar_ocr_b200 import models
    return models.get_blob("rec")


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
