import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "attribute-guided-image-generation-from-layout_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # on a GPU box the CUDA library must exist before the first kernel call (the product has no fallback): build it if
    # the in-tree .so did not travel with the sources
    try:
        import torch
        if torch.cuda.is_available():
            import __graft_entry__ as ge
            if not os.path.exists(ge.LIB):
                ge.build()
    except Exception as e:          # the tests themselves will report a missing library loudly
        print("conftest: could not build libb200gan.so: %s" % e, file=sys.stderr)


@pytest.fixture
def emul(monkeypatch):
    """Replace the kernel namespace with the CPU emulation of the C ABI (host-wiring tests only)."""
    from b200gan import _lib, ops
    from abi_emul import EmulKernels
    k = EmulKernels()
    monkeypatch.setattr(_lib, "K", k)
    ops._PLANS.clear()
    ops._LINSPACE.clear()
    yield k
    ops._PLANS.clear()
    ops._LINSPACE.clear()
