import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def r3d():
    """The product package (its name starts with a digit, hence importlib)."""
    return importlib.import_module("3d_recognizer_b200")


@pytest.fixture(scope="session")
def ops(r3d):
    return importlib.import_module("3d_recognizer_b200.ops")


@pytest.fixture(scope="session")
def built_lib(r3d):
    """Builds lib/libr3d_b200.so if nvcc is here and it is stale (cross-compiles without a GPU)."""
    import shutil
    build = importlib.import_module("3d_recognizer_b200.build")
    if shutil.which("nvcc"):
        build.build_library()
    return build.LIB_PATH


@pytest.fixture(scope="session")
def oracle_built():
    from oracle import knn as oknn
    oknn.build(ref=True)
    return oknn
