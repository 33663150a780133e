"""ctypes binding of libdgp_b200.so (the C-ABI declared in include/dgp_b200.h).

There is no fallback: if the shared library is missing, or CUDA is unavailable when a compute entry point is
called, an exception is raised.  PyTorch only provides device memory and the current stream.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdgp_b200.so")


def use_library(path):
    """Load another build of the library instead (tools/kbench.py: phase-counter and test-hook builds).  Must be called
    before the first lib(); the product never calls it and reads no environment variable."""
    global LIB_PATH
    if _lib is not None:
        raise DgpError("use_library() must be called before the library is loaded")
    LIB_PATH = path

DGP_OK, DGP_ERR_ARG, DGP_ERR_LAUNCH, DGP_ERR_WORKSPACE, DGP_ERR_UNSUPPORTED = 0, -1, -2, -3, -4
DGP_KIND_RBF = 0
DGP_KIND_MATERN52 = 1
DGP_MAX_TASKS = 16
DGP_MAX_D = 64
DGP_MAX_R = 64


class CondDesc(ctypes.Structure):
    """struct dgp_cond_desc (include/dgp_b200.h)."""
    _fields_ = [
        ("M", ctypes.c_int32), ("D", ctypes.c_int32), ("Dx", ctypes.c_int32), ("R", ctypes.c_int32),
        ("ldr", ctypes.c_int32), ("col0", ctypes.c_int32), ("T", ctypes.c_int32),
        ("kind", ctypes.c_int32 * DGP_MAX_TASKS),
        ("jitter", ctypes.c_double),
        ("Z", ctypes.c_void_p), ("theta", ctypes.c_void_p), ("q_mu", ctypes.c_void_p), ("q_sqrt", ctypes.c_void_p),
        ("mean_A", ctypes.c_void_p), ("mean_b", ctypes.c_void_p),
    ]


_P = ctypes.POINTER(CondDesc)
_vp, _i32, _i64, _u64, _sz, _dbl = (ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64, ctypes.c_size_t,
                                    ctypes.c_double)

# name -> (restype, argtypes); every symbol include/dgp_b200.h declares
SIGNATURES = {
    "dgp_cond_ws_bytes": (_sz, [_P, _i64, ctypes.c_int]),
    "dgp_cond_prepare": (ctypes.c_int, [_P, _vp, _sz, ctypes.c_int, _vp, _vp]),
    "dgp_cond_status": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(_i32)]),
    "dgp_cond_status_async": (ctypes.c_int, [_vp, _vp, _vp]),
    "dgp_cond_forward": (ctypes.c_int, [_P, _vp, _vp, _i64, _i64, _vp, _u64, _vp, _u64, _i64, _i64, _vp, _vp, _vp, _i32,
                                        _vp, _vp, _vp, _vp]),
    "dgp_usave_bytes": (_sz, [_P, _i64]),
    "dgp_cond_forward_group": (ctypes.c_int, [_i32, _vp, _vp, _vp, _i64, _i64, _vp, _u64, _vp, _u64, _i64, _i64, _vp, _vp,
                                              _vp, _i32, _vp, _vp, _vp, _vp]),
    "dgp_cond_backward_group": (ctypes.c_int, [_i32, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp,
                                               _vp, _vp, _i32, _i32, _vp]),
    "dgp_sample_expand": (ctypes.c_int, [_vp, _vp, _vp, _u64, _vp, _u64, _i64, _i32, _i64, _i32, _i32, _vp, _i32, _vp,
                                         _i32, _vp]),
    "dgp_sample_reduce": (ctypes.c_int, [_vp, _i32, _vp, _i32, _vp, _vp, _i32, _i64, _i32, _i32, _vp, _vp, _vp]),
    "dgp_lik_gaussian": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i64, _i32, _vp, _i32, _dbl, _vp, _vp, _vp, _vp,
                                        _vp, _sz, _vp]),
    "dgp_lik_ws_bytes": (_sz, [_i64, _i32]),
    "dgp_cond_backward": (ctypes.c_int, [_P, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp,
                                         _i32, _vp, _i32, _i32, _vp]),
    "dgp_cond_backward_gram": (ctypes.c_int, [_P, _vp, _vp, _i64, _i32, _vp]),
    "dgp_cond_backward_params": (ctypes.c_int, [_P, _vp, _dbl, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dgp_cond_fullcov_ws_bytes": (_sz, [_P, _i64]),
    "dgp_cond_fullcov": (ctypes.c_int, [_P, _vp, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "dgp_chol_sample_ws_bytes": (_sz, [_i64, _i32]),
    "dgp_chol_sample": (ctypes.c_int, [_vp, _i64, _i64, _i64, _i64, _i32, _dbl, _vp, _i32, _vp, _i64, _i64, _vp, _vp, _sz,
                                       _vp, _vp]),
    "dgp_pca_ws_bytes": (_sz, [_i64, _i32]),
    "dgp_pca_weights": (ctypes.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _sz, _vp, _vp, _vp]),
    "dgp_transform_params": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i32, _dbl, _vp, _vp]),
    "dgp_adam_step": (ctypes.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _dbl, _dbl, _dbl, _dbl, _vp, _vp]),
    "dgp_set_u64": (ctypes.c_int, [_vp, _u64, _vp]),
    "dgp_finish_elbo": (ctypes.c_int, [_vp, _vp, _i32, _dbl, _vp]),
    "dgp_zero": (ctypes.c_int, [_vp, _sz, _vp]),
    "dgp_kernel_K": (ctypes.c_int, [_P, _vp, _i64, _vp, _i64, _i32, _vp, _vp]),
    "dgp_kernel_Kdiag": (ctypes.c_int, [_P, _vp, _i64, _vp, _vp]),
    "dgp_mpad": (_i32, [_i32]),
    "dgp_philox_normal": (ctypes.c_int, [_vp, _i64, _i32, _i32, _u64, _u64, _i64, _i64, _vp]),
    "dgp_strerror": (ctypes.c_char_p, [ctypes.c_int]),
    "dgp_version": (ctypes.c_char_p, []),
    "dgp_launch_count": (_i64, []),
}

_lib = None


class DgpError(RuntimeError):
    pass


def lib():
    """Load libdgp_b200.so once.  Raises if it has not been built (python __graft_entry__.py / make -C csrc)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DgpError(f"{LIB_PATH} not found: build it with `make -C dgplib_b200/csrc` "
                           "(there is no CPU or PyTorch fallback for the DSDGP hot path)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().dgp_strerror(rc).decode()
        raise DgpError(f"{what} failed: {msg} (code {rc})")


def require_cuda():
    if not torch.cuda.is_available():
        raise DgpError("dgplib_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def ptr(t):
    """Device pointer of a contiguous float64 CUDA tensor (or None)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    return ctypes.c_void_p(t.data_ptr())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device (the raw-handle query when this torch has it: building
    a torch.cuda.Stream object per launch was a visible part of the per-step host time of small models)."""
    if _raw_stream is not None:
        return ctypes.c_void_p(_raw_stream(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def mpad(M):
    return (int(M) + 63) // 64 * 64


def launch_count():
    return int(lib().dgp_launch_count())
