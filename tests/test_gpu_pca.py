"""GPU PCA behind find_weights (dgplib/layers.py:97-121) against the reference's golden vector
(tests/test_layer.py:33-38) and the oracle's LAPACK-SVD restatement.  Singular vectors are compared after aligning the
per-vector sign, which an SVD leaves open."""
import ctypes

import numpy as np
import pytest
import torch

from dgplib_b200 import _cabi as C, settings
from dgplib_b200.engine import pca_weights
from dgplib_b200.kernels import RBF
from dgplib_b200.layers import InputLayer, find_weights
from oracle import dsdgp_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _data(n, d, seed):
    rng = np.random.RandomState(seed)
    Q, _ = np.linalg.qr(rng.randn(d, d))
    return (rng.randn(n, d) * np.linspace(3.0, 0.5, d)) @ Q


def _align(W, Wref):
    return W * np.sign(np.sum(W * Wref, axis=0, keepdims=True))


def test_reference_golden_vector(cuda):
    X = np.array([[2, 4], [1, 3], [0, 0], [0, 0]], dtype=np.float64)
    W = find_weights(2, 1, X, backend="cuda")
    assert W.shape == (2, 1)
    assert np.allclose(W, np.array([-0.4, -0.91]).reshape(2, 1), atol=1e-2)      # the reference's own assertion
    assert np.allclose(_align(W, O.find_weights(2, 1, X)), O.find_weights(2, 1, X), atol=1e-12)
    assert np.allclose(find_weights(2, 2, X, backend="cuda"), np.eye(2))
    assert np.allclose(find_weights(1, 2, X[:, :1], backend="cuda"), np.array([[1.0, 0.0]]))


@pytest.mark.parametrize("n,d,r", [(10, 5, 2), (1000, 8, 3), (5003, 16, 15), (20000, 33, 5), (777, 64, 10), (3, 2, 1),
                                   (100000, 16, 8), (4099, 7, 6)])
def test_matches_lapack_svd(cuda, n, d, r):
    X = _data(n, d, n + d)
    W, sing = pca_weights(X, r, return_singular_values=True)
    Wref = O.find_weights(d, r, X)
    assert W.shape == (d, r)
    assert np.max(np.abs(_align(W, Wref) - Wref)) < TOL
    sref = np.linalg.svd(X, compute_uv=False)
    assert np.max(np.abs(sing[:len(sref)] - sref) / sref[0]) < 1e-11
    assert np.max(np.abs(W.T @ W - np.eye(r))) < 1e-12                             # orthonormal columns
    big = W[np.argmax(np.abs(W), axis=0), np.arange(r)]
    assert np.all(big < 0)                                                          # documented sign convention


def test_multitask_drops_task_column(cuda):
    rng = np.random.RandomState(5)
    X = np.hstack([_data(500, 6, 1), rng.randint(0, 3, (500, 1)).astype(np.float64)])
    W = find_weights(6, 2, X, multitask=True, backend="cuda")
    Wref = O.find_weights(6, 2, X, multitask=True)
    assert W.shape == (7, 2) and np.all(W[-1] == 0.0)
    assert np.max(np.abs(_align(W, Wref) - Wref)) < TOL


def test_row_pitch_and_deterministic(cuda):
    """C-ABI call on a strided view (ldx > D), twice: bit-identical output (fixed-order reductions)."""
    lib = C.lib()
    X = torch.tensor(np.hstack([_data(3000, 12, 2), np.full((3000, 3), 7.0)]), device="cuda")   # 3 columns to ignore
    n, D, R = 3000, 12, 4
    wsb = lib.dgp_pca_ws_bytes(n, D)
    outs = []
    for _ in range(2):
        ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
        W = torch.zeros(D, R, dtype=torch.float64, device="cuda")
        C.check(lib.dgp_pca_weights(C.ptr(X), n, X.shape[1], D, R, C.ptr(ws), wsb, C.ptr(W), None, C.stream_ptr()), "pca")
        outs.append(W.cpu().numpy())
    assert np.array_equal(outs[0], outs[1])
    Wref = O.find_weights(D, R, X[:, :D].cpu().numpy())
    assert np.max(np.abs(_align(outs[0], Wref) - Wref)) < TOL
    assert lib.dgp_pca_weights(C.ptr(X), n, 8, D, R, C.ptr(ws), wsb, C.ptr(W), None, C.stream_ptr()) == C.DGP_ERR_ARG
    assert lib.dgp_pca_weights(C.ptr(X), n, 15, D, R, C.ptr(ws), 16, C.ptr(W), None, C.stream_ptr()) == C.DGP_ERR_WORKSPACE


def test_initialize_forward_with_cuda_backend(cuda, monkeypatch):
    X, Z = _data(400, 5, 3), _data(20, 5, 4)
    outs = {}
    for backend in ("numpy", "cuda"):
        monkeypatch.setattr(settings, "pca_backend", backend)
        layer = InputLayer(5, 2, 20, RBF(5))
        outs[backend] = layer.initialize_forward(X, Z)
    sgn = np.sign(np.sum(outs["numpy"][0] * outs["cuda"][0], axis=0, keepdims=True))
    assert np.max(np.abs(outs["numpy"][0] - sgn * outs["cuda"][0])) < 1e-8
    assert np.max(np.abs(outs["numpy"][1] - sgn * outs["cuda"][1])) < 1e-8
