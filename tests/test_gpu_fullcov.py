"""Full-covariance branch against the oracle and the reference's own shape tests
(tests/test_dsdgp.py:99-141, tests/test_multikernel_layer.py:67-116)."""
import numpy as np
import pytest
import torch

import dgplib_b200 as dg
from dgplib_b200.cascade import Sequential
from dgplib_b200.kernels import RBF, Matern52, White
from dgplib_b200.layers import InputLayer, OutputLayer
from dgplib_b200.likelihoods import Gaussian
from dgplib_b200.mean_functions import Linear
from dgplib_b200.multikernel_layers import MultikernelInputLayer
from oracle import dsdgp_oracle as O
from _models import build_dsdgp, oracle_layers, randomize_q, rel_err, t

pytestmark = pytest.mark.gpu
TOL = 1e-9


def test_reference_full_cov_shapes(cuda):
    rng = np.random.RandomState(42)
    X, Y, Z, Xs = rng.randn(100, 2), rng.randn(100, 1), rng.randn(10, 2), rng.randn(10, 2)
    il = InputLayer(2, 1, 10, RBF(2) + White(2), mean_function=Linear(A=np.ones((2, 1))))
    ol = OutputLayer(1, 1, 10, RBF(1) + White(1))
    model = dg.DSDGP(X, Y, Z, Sequential([il, ol]), Gaussian())
    model.compile()
    mu, sigma = model.predict_f_full_cov(Xs, 1)
    assert mu.shape == (1, 10, 1) and sigma.shape == (1, 10, 10, 1) and np.all(sigma > -1e-6)
    fs, fmeans, fvars = model.predict_all_layers_full_cov(Xs, 1)
    for f, m, v in zip(fs, fmeans, fvars):
        assert f.shape == (1, 10, 1) and m.shape == (1, 10, 1) and v.shape == (1, 10, 10, 1) and np.all(v > -1e-6)
    assert model.predict_f_samples(Xs, 10).shape == (10, 10, 1)


def test_multikernel_full_cov_shapes(cuda):
    rng = np.random.RandomState(42)
    for out_dim in (3, 9):
        layer = MultikernelInputLayer(2, out_dim, 5, [RBF(2), RBF(2), RBF(2)], share_Z=False)
        layer.initialize_forward(rng.randn(10, 2), rng.randn(5, 2))
        m, v = layer._build_predict(rng.randn(4, 10, 2), full_cov=True, stochastic=True)
        assert m.shape == (4, 10, out_dim) and v.shape == (4, 10, 10, out_dim)
        m, v = layer._build_predict(rng.randn(10, 2), full_cov=True, stochastic=False)
        assert m.shape == (10, out_dim) and v.shape == (out_dim, 10, 10)


def test_full_cov_matches_oracle(cuda):
    model, X, Y = build_dsdgp(L=3, M=24, D=3, N=120, S=2, kind="matern52", ard=True, white=1e-3)
    model.compile()
    rng = np.random.RandomState(5)
    N, S = 33, 3
    Xs = rng.randn(N, 3)
    dims = [l.output_dim for l in model.layers.layers]
    z = [rng.randn(S, R, N) for R in dims]
    ol = oracle_layers(model.layers.layers)
    oF, oM, oV = O.propagate_full_cov(ol, t(Xs), S, [t(a) for a in z])
    F, M, V = model.predict_all_layers_full_cov(Xs, S, eps=z)
    for l in range(3):
        assert rel_err(M[l], oM[l]) < TOL and rel_err(V[l], oV[l]) < TOL and rel_err(F[l], oF[l]) < 1e-8
    mu, var = model.predict_f_full_cov(Xs, S, eps=z)
    assert rel_err(mu, oM[-1]) < TOL and rel_err(var, oV[-1]) < TOL
    # the diagonal of the full covariance is the marginal variance of the diag path on the same inputs
    m1, v1 = model.layers.layers[0]._build_predict(Xs[None], full_cov=True)
    m2, v2 = model.layers.layers[0]._build_predict(Xs[None], full_cov=False)
    assert rel_err(torch.diagonal(v1[0], dim1=0, dim2=1).t(), v2[0]) < 1e-11 and rel_err(m1, m2) < 1e-13
    # predict_f_samples with injected draws
    z1 = [rng.randn(1, R, N) for R in dims]
    Vd = rng.randn(1, N, 7)
    got = model.predict_f_samples(Xs, 7, eps=z1, V=Vd)
    want = O.predict_f_samples(ol, t(Xs), 7, [t(a) for a in z1], t(Vd))
    assert got.shape == (7, N, 1) and rel_err(got, want) < 1e-8


def test_normal_sample_full_cov_util(cuda):
    from dgplib_b200.utils import normal_sample
    rng = np.random.RandomState(1)
    S, N, D = 2, 17, 3
    Bm = rng.randn(S, D, N, N)
    var = torch.tensor(np.einsum("sdij,sdkj->sikd", Bm, Bm)).cuda()          # SPD per (s, d), laid out [S,N,N,D]
    mean = torch.tensor(rng.randn(S, N, D)).cuda()
    z = torch.tensor(rng.randn(S, D, N)).cuda()
    got = normal_sample(mean, var, full_cov=True, eps=z)
    want = O.normal_sample_full_cov(mean.cpu(), var.cpu(), z.cpu())
    assert got.shape == (S, N, D) and rel_err(got, want) < 1e-10


@pytest.mark.parametrize("N", [1, 64, 130, 1024, 1100, 1500])
def test_full_cov_sizes(cuda, N):
    """Single points, exact tile multiples, ragged sizes, the largest size the one-CTA Cholesky takes (1024) and sizes
    above it (blocked multi-launch factorisation; 1100 pads to 1152 = four full 256-blocks and one of 128)."""
    model, X, Y = build_dsdgp(L=2, M=20, D=2, N=60, S=1, white=1e-3)
    model.compile()
    rng = np.random.RandomState(N)
    Xs = rng.randn(N, 2)
    dims = [l.output_dim for l in model.layers.layers]
    z = [rng.randn(1, R, N) for R in dims]
    ol = oracle_layers(model.layers.layers)
    oF, oM, oV = O.propagate_full_cov(ol, t(Xs), 1, [t(a) for a in z])
    F, M, V = model.predict_all_layers_full_cov(Xs, 1, eps=z)
    for l in range(2):
        assert M[l].shape == tuple(oM[l].shape) and V[l].shape == tuple(oV[l].shape)
        assert rel_err(M[l], oM[l]) < TOL and rel_err(V[l], oV[l]) < TOL and rel_err(F[l], oF[l]) < 1e-7


def test_full_cov_refuses_more_than_16384_points(cuda):
    from dgplib_b200 import _cabi
    assert _cabi.lib().dgp_chol_sample_ws_bytes(16385, 1) == 0 and _cabi.lib().dgp_chol_sample_ws_bytes(16384, 1) > 0
    assert _cabi.lib().dgp_chol_sample_ws_bytes(2048, 129) == 0


def test_multikernel_full_cov_matches_oracle(cuda):
    rng = np.random.RandomState(3)
    layer = MultikernelInputLayer(2, 4, 12, [RBF(2, lengthscales=0.9) + White(2, 1e-3), Matern52(2, lengthscales=1.4)],
                                  share_Z=False, mean_function=Linear(A=rng.randn(2, 4)))
    layer.initialize_forward(rng.randn(30, 2), rng.randn(12, 2))
    layer.q_mu.assign(0.2 * rng.randn(12, 4))
    layer.q_sqrt.assign(np.eye(12)[None] + 0.1 * np.tril(rng.randn(4, 12, 12)))
    Xs = rng.randn(2, 21, 2)
    m, v = layer._build_predict(Xs, full_cov=True)
    om, ov = O.layer_build_predict_full_cov(oracle_layers([layer])[0], t(Xs))
    assert m.shape == (2, 21, 4) and v.shape == (2, 21, 21, 4)
    assert rel_err(m, om) < TOL and rel_err(v, ov) < TOL
