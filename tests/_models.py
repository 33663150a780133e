"""Test helpers: build a dgplib_b200 model and the equivalent plain-dict description the oracle consumes.
The oracle side shares no code with the product: it only receives parameter VALUES copied out of the model."""
import numpy as np
import torch

import dgplib_b200 as dg
from dgplib_b200.cascade import Sequential, MultitaskSequential
from dgplib_b200.kernels import RBF, Matern52, White, Stationary, Sum
from dgplib_b200.layers import InputLayer, HiddenLayer, OutputLayer
from dgplib_b200.likelihoods import Gaussian, SwitchedLikelihood
from dgplib_b200.mean_functions import Linear
from dgplib_b200.multikernel_layers import MultikernelInputLayer, MultikernelHiddenLayer, MultikernelOutputLayer
from dgplib_b200.specialized_kernels import SwitchedKernel

T64 = torch.float64


def t(x):
    return torch.tensor(np.array(x, dtype=np.float64), dtype=T64)


from oracle.bridge import oracle_kernel, oracle_layers, oracle_lik  # noqa: E402,F401


def randomize_q(model, seed=1, scale=0.1):
    """SURVEY.md 8d parity parameters: q_mu = 0.1 N(0,1), q_sqrt = I + 0.1 tril(N(0,1)) (+ junk above the diagonal,
    which band_part must ignore)."""
    rs = np.random.RandomState(seed)
    for layer in model.layers.layers:
        M, R = layer.q_mu.shape
        layer.q_mu.assign(scale * rs.randn(M, R))
        q = np.eye(M)[None] + scale * np.tril(rs.randn(R, M, M)) + 0.3 * np.triu(rs.randn(R, M, M), 1)
        layer.q_sqrt.assign(q)


def make_eps(model, S, B, seed=2):
    return [np.random.RandomState(seed + l).randn(S, B, layer.output_dim)
            for l, layer in enumerate(model.layers.layers)]


def make_kernel(kind, D, ard=False, white=1e-5, ls_scale=1.0, variance=1.0):
    cls = RBF if kind == "rbf" else Matern52
    ls = np.sqrt(D) * ls_scale
    if ard:
        ls = ls * (0.8 + 0.4 * np.random.RandomState(7).rand(D))
    k = cls(D, variance=variance, lengthscales=ls, ARD=ard)
    return k + White(D, variance=white) if white else k


def synthetic_data(N, D, M, seed=0, D_Y=1):
    """SURVEY.md 8d synthetic inputs."""
    rng = np.random.RandomState(seed)
    X = rng.randn(N, D)
    Y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.randn(N, D_Y)
    Z = X[rng.permutation(N)[:M]].copy()
    return X, Y, Z


def build_dsdgp(L=3, M=32, D=4, N=200, S=3, kind="rbf", ard=False, minibatch=None, hidden=None, seed=0, white=1e-5,
                lik_var=0.1, linear=True):
    X, Y, Z = synthetic_data(N, D, M, seed)
    hidden = D if hidden is None else hidden
    dims = [D] + [hidden] * (L - 1) + [1]
    layers = []
    for l in range(L):
        cls = InputLayer if l == 0 else (OutputLayer if l == L - 1 else HiddenLayer)
        if L == 1:
            cls = InputLayer
        mf = Linear(A=np.zeros((dims[l], dims[l + 1]))) if (linear and l < L - 1) else None
        layers.append(cls(dims[l], dims[l + 1], M, make_kernel(kind, dims[l], ard, white), mean_function=mf))
    model = dg.DSDGP(X, Y, Z, Sequential(layers), Gaussian(variance=lik_var), minibatch_size=minibatch, num_samples=S)
    randomize_q(model)
    return model, X, Y


def oracle_grads_by_param(model, glist):
    """Map the engine's flat grad list back to names comparable with oracle.leaves()."""
    out, it = {}, iter(glist)
    for l, layer in enumerate(model.layers.layers):
        for g, _ in enumerate(layer.conditionals()):
            out[f"l{l}.g{g}.Z"] = next(it)
            out[f"l{l}.g{g}.theta"] = next(it)
        out[f"l{l}.q_mu"] = next(it)
        out[f"l{l}.q_sqrt"] = next(it)
        mf = layer.mean_function
        if type(mf).__name__ == "Linear" and (mf.A.trainable or mf.b.trainable):
            out[f"l{l}.mean.A"] = next(it)
            out[f"l{l}.mean.b"] = next(it)
    out["lik"] = next(it)
    return out


def oracle_theta_grad(og, l, g, kern, D):
    """Assemble the oracle's gradient in the engine's theta layout [T, 2+D]."""
    def one(prefix, k):
        ls = og[f"{prefix}.lengthscales"]
        ls = ls.reshape(-1)
        if ls.numel() == 1:
            # scalar lengthscale: the engine differentiates w.r.t. the D expanded copies; compare sums instead
            ls = None
        w = og.get(f"{prefix}.white", torch.zeros((), dtype=T64)).reshape(1)
        return og[f"{prefix}.variance"].reshape(1), w, ls
    if kern["kind"] == "switched":
        return [one(f"l{l}.g{g}.kern.k{t}", kk) for t, kk in enumerate(kern["kernels"])]
    return [one(f"l{l}.g{g}.kern", kern)]


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=T64).cpu()
    b = torch.as_tensor(b, dtype=T64).cpu()
    denom = float(b.abs().max())
    return float((a - b).abs().max()) / (denom if denom > 0 else 1.0)
