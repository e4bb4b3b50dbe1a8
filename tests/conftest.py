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
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _native_library():
    """libvilco_b200.so is a build artefact (git-ignored): build it once per session when it is missing, so that any subset of
    the tests can be run from a fresh checkout.  (The product itself never builds on demand: vilco_b200.lib raises.)"""
    from vilco_b200 import lib
    if not os.path.exists(lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    yield


@pytest.fixture(autouse=True)
def _precision_mode(request):
    """Every GPU test runs in the SHIPPED default operand-format policy ("mixed", vilco_b200/ops.py) unless its module sets
    PRECISION = "fp16x3" (operator-level tests whose tolerances state the exact split-operand arithmetic)."""
    if "gpu" not in request.keywords:
        yield
        return
    from vilco_b200 import ops
    want = getattr(request.module, "PRECISION", "mixed")
    prev = ops.precision()
    ops.set_precision(want)
    yield
    ops.set_precision(prev)
