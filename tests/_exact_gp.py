"""Exact-GP anchor shared by the CPU and GPU parity tests: scikit-learn (an independent third-party implementation)
provides kernel matrices and the exact log marginal likelihood that a single-layer model with Z = X and the optimal
q(u) must reproduce."""
import numpy as np


def sk_kernel(kind, variance, ls):
    from sklearn.gaussian_process import kernels as sk
    base = sk.RBF(length_scale=ls) if kind == "rbf" else sk.Matern(length_scale=ls, nu=2.5)
    return sk.ConstantKernel(variance) * base


def exact_gp_case(kind, seed=0, N=24, D=2):
    """Inputs on a jittered grid (well-conditioned K), the optimal whitened q(u) for Z = X, and scikit-learn's exact-GP
    log marginal likelihood: with inducing points at the data and the optimal q the collapsed bound is tight, so the
    whole single-layer ELBO (conditional + KL + Gaussian expectations) must reproduce an independent library's number."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    rng = np.random.RandomState(seed)
    g = np.stack(np.meshgrid(*[np.arange(5.0)] * D), -1).reshape(-1, D)[:N]
    X = 1.4 * g + 0.2 * rng.rand(N, D)
    Y = np.sin(X.sum(1, keepdims=True)) + 0.3 * rng.randn(N, 1)
    variance, ls, noise = 1.7, np.array([1.1, 0.8])[:D], 0.25
    skk = sk_kernel(kind, variance, ls)
    lml = GaussianProcessRegressor(kernel=skk, alpha=noise, optimizer=None).fit(X, Y).log_marginal_likelihood()
    # the same model with the noise as a WhiteKernel hyper-parameter, for d lml / d (variance, lengthscales, noise):
    # by the envelope theorem these equal the ELBO's partial derivatives at the optimal q (where d ELBO / d q = 0)
    from sklearn.gaussian_process.kernels import WhiteKernel
    gpr = GaussianProcessRegressor(kernel=skk + WhiteKernel(noise), alpha=0.0, optimizer=None).fit(X, Y)
    lml2, glog = gpr.log_marginal_likelihood(gpr.kernel_.theta, eval_gradient=True)
    assert abs(lml2 - lml) < 1e-10 * abs(lml)
    names = [h.name for h in gpr.kernel_.hyperparameters]
    assert names == ["k1__k1__constant_value", "k1__k2__length_scale", "k2__noise_level"], names
    gtheta = glog / np.exp(gpr.kernel_.theta)          # scikit-learn differentiates w.r.t. log(theta)
    grad = {"variance": gtheta[0], "lengthscales": gtheta[1:1 + D], "noise": gtheta[1 + D]}
    K = skk(X)
    L = np.linalg.cholesky(K + 1e-12 * np.eye(N))
    B = np.linalg.solve(K + noise * np.eye(N), K)
    m_f, S_f = B.T @ Y, K - K @ B                      # exact posterior over f(X)
    Li = np.linalg.inv(L)
    S_v = Li @ S_f @ Li.T
    q_mu, q_sqrt = Li @ m_f, np.linalg.cholesky(0.5 * (S_v + S_v.T))[None]
    return dict(X=X, Y=Y, variance=variance, ls=ls, noise=noise, q_mu=q_mu, q_sqrt=q_sqrt, lml=float(lml), K=K, kernel=skk, grad=grad)
