import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session", autouse=True)
def _built_library(request):
    """A fresh checkout has no built library (it is git-ignored): build it once, exactly as __graft_entry__.build() does --
    but only for sessions that collected a test which loads it (the oracle / host-logic tests run on machines without
    nvcc)."""
    needs = any(("gpu" in item.keywords) or item.fspath.basename in ("test_cabi.py",) for item in request.session.items)
    if needs and not os.path.exists(os.path.join(ROOT, "dgplib_b200", "libdgp_b200.so")):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def hooked_lib():
    """libdgp_b200_test.so: the same sources built with -DDGP_TEST_HOOKS (`make test_lib`), whose extra entry points let a
    test pin every row-tile instantiation.  Loaded beside the product library; nothing in dgplib_b200/ knows about it."""
    import ctypes
    import subprocess
    from dgplib_b200 import _cabi
    path = os.path.join(ROOT, "dgplib_b200", "libdgp_b200_test.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-C", os.path.join(ROOT, "dgplib_b200", "csrc"), "test_lib", "-j8"], check=True,
                       stdout=subprocess.DEVNULL)
    handle = ctypes.CDLL(path)
    for name, (res, args) in _cabi.SIGNATURES.items():
        fn = getattr(handle, name)
        fn.restype, fn.argtypes = res, args
    handle.dgp_test_force_tiles.argtypes = [ctypes.c_int, ctypes.c_int]
    handle.dgp_test_force_tiles.restype = None
    yield handle
    handle.dgp_test_force_tiles(0, 0)


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("this test is marked gpu and needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda:0")
