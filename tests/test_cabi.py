"""The C-ABI library loads and exports every symbol include/dgp_b200.h declares (no GPU, no compute calls)."""
import ctypes
import os
import re

from dgplib_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dgp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dgp_[A-Za-z_0-9]+)\s*\(", text)))


def test_header_declares_the_binding_table():
    assert set(declared_symbols()) == set(_cabi.SIGNATURES), (declared_symbols(), sorted(_cabi.SIGNATURES))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_cabi.LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    handle = ctypes.CDLL(_cabi.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(handle, name), name
    _cabi.lib()


def test_host_side_entry_points():
    lib = _cabi.lib()
    assert lib.dgp_mpad(1) == 64 and lib.dgp_mpad(64) == 64 and lib.dgp_mpad(65) == 128 and lib.dgp_mpad(500) == 512
    assert b"sm_100a" in lib.dgp_version()
    assert lib.dgp_strerror(0) == b"ok" and lib.dgp_strerror(-3) == b"workspace too small"
    assert _cabi.mpad(300) == lib.dgp_mpad(300)


def test_desc_struct_layout_and_workspace_size():
    d = _cabi.CondDesc()
    assert ctypes.sizeof(d) == 7 * 4 + 16 * 4 + 4 + 8 + 6 * 8      # 7 ints, kind[16], pad, jitter, 6 pointers
    d.M, d.D, d.Dx, d.R, d.ldr, d.col0, d.T = 256, 8, 8, 8, 8, 0, 1
    lib = _cabi.lib()
    fwd = lib.dgp_cond_ws_bytes(ctypes.byref(d), 1000, 0)
    bwd = lib.dgp_cond_ws_bytes(ctypes.byref(d), 1000, 1)
    assert fwd > 10 * 256 * 256 * 8 and bwd > fwd + 1000 * 256 * 8
    bad = _cabi.CondDesc()
    assert lib.dgp_cond_ws_bytes(ctypes.byref(bad), 10, 0) == 0


def test_argument_validation_without_a_gpu():
    lib = _cabi.lib()
    d = _cabi.CondDesc()
    d.M, d.D, d.Dx, d.R, d.ldr, d.col0, d.T = 2000, 8, 8, 8, 8, 0, 1     # M too large, null pointers
    assert lib.dgp_cond_prepare(ctypes.byref(d), None, 0, 0, None, None) < 0
    assert lib.dgp_cond_forward(ctypes.byref(d), None, None, 1, 1, None, 0, None, 0, 0, 0, None, None, None, 8, None, None, None, None) < 0
    assert lib.dgp_lik_gaussian(None, None, None, 1, 1, 1, 1, None, 1, 1.0, None, None, None, None, None, 0, None) < 0
    assert lib.dgp_philox_normal(None, 1, 1, 1, 0, 0, 0, 0, None) < 0


def test_no_cpu_fallback():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_cabi.DgpError):
        _cabi.require_cuda()


def test_product_and_tools_never_touch_the_oracle():
    """oracle/ is test infrastructure: nothing shipped (the package, the C sources, tools/) may import or mention it as a
    code path; bench.py and __graft_entry__.py may, as the checker / timed CPU baseline only."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    offenders = []
    for sub in ("dgplib_b200", "tools", "include"):
        for dirpath, _, files in os.walk(os.path.join(root, sub)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    text = open(os.path.join(dirpath, f), encoding="utf-8", errors="replace").read()
                    if pat.search(text):
                        offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders
