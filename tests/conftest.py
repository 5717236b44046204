import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Both in-tree libraries (libevc.so: nvcc cross-compiles without a GPU; libevc_reader.so: g++) exist before
    any test imports the package; a no-op when they are newer than their sources."""
    import __graft_entry__ as g
    try:
        g.build()
    except Exception:                     # e.g. no compiler on this machine: prebuilt in-tree libraries still serve
        if not (os.path.exists(g.LIB) and os.path.exists(g.READER_LIB)):
            raise


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
