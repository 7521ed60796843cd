import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle

    return oracle.lib()


@pytest.fixture(scope="session")
def wdb_lib():
    """libwdb200.so, built in-tree if the sources are newer (nvcc cross-compiles on CPU)."""
    from warp_drive_b200 import build as wbuild
    from warp_drive_b200 import lib as wlib

    if wbuild.needs_build():
        wbuild.build()
    return wlib.load()


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))
