"""Op-for-op float64 restatement (PyTorch-CPU) of the reference DSDGP graph.  TEST INFRASTRUCTURE ONLY
(see ``oracle/__init__.py``).  ELBO / gradient / prediction values: parity unpinned by the reference's own tests.

The restatement keeps the reference's inefficiencies on purpose (S-tiled layer-0 input, dense matmul on
``tril(q_sqrt)``, materialised ``[R, M, N']`` LTA tensor, autograd backward) because it is also the timed CPU
baseline ("CPU restatement of the reference path"; TF 1.8 / GPflow 1.2 are not installable here).

Model description (plain dicts of torch float64 tensors; nothing is shared with ``dgplib_b200``):

    kern  = {"kind": "rbf"|"matern52"|"const" (test double), "input_dim": D, "variance": t[], "lengthscales": t[] or t[D],
             "white": t[] or None}
          | {"kind": "switched", "kernels": [kern, ...]}           # dgplib/specialized_kernels.py
    layer = {"groups": [{"kernel": kern, "Z": t[M, Dz]}, ...],    # 1 group = Layer; K groups = MultikernelLayer
             "q_mu": t[M, R], "q_sqrt": t[R, M, M],
             "mean": None | {"A": t[Dz, R], "b": t[R]}}
    lik   = {"kind": "gaussian", "variance": t[]} | {"kind": "switched", "variances": [t[], ...]}
"""
import math

import torch

JITTER = 1e-6  # gpflow.settings.numerics.jitter_level, read at dgplib/utils.py:12


# ----------------------------------------------------------------------------------------------- kernels
def _slice(X, input_dim):
    # gpflow Kernel._slice with the default active_dims = slice(input_dim): trailing columns are ignored.
    return X[:, :input_dim]


def _scaled_sq_dist(X, X2, ls):
    # gpflow Stationary.scaled_square_dist: norms + (-2 X X2^T); no clamp at zero.
    Xs = X / ls
    X2s = X2 / ls
    return -2.0 * Xs @ X2s.t() + (Xs * Xs).sum(1)[:, None] + (X2s * X2s).sum(1)[None, :]


def _stationary_K(kern, X, X2):
    if kern["kind"] == "const":
        # the reference's own test double: DummyKernel of tests/test_specialized_kernels.py:13-26, K = val * ones
        return kern["variance"] * torch.ones(X.shape[0], X2.shape[0], dtype=X.dtype)
    r2 = _scaled_sq_dist(X, X2, kern["lengthscales"])
    if kern["kind"] == "rbf":
        return kern["variance"] * torch.exp(-r2 / 2.0)
    if kern["kind"] == "matern52":
        r = torch.sqrt(r2 + 1e-12)
        s5 = math.sqrt(5.0)
        return kern["variance"] * (1.0 + s5 * r + 5.0 / 3.0 * r * r) * torch.exp(-s5 * r)
    raise ValueError(kern["kind"])


def kern_K(kern, X, X2=None):
    """K(X, X2); X2=None means the symmetric call K(X) (White contributes only there)."""
    if kern["kind"] == "switched":
        return switched_K(kern, X, X2)
    D = kern["input_dim"]
    Xa = _slice(X, D)
    X2a = Xa if X2 is None else _slice(X2, D)
    K = _stationary_K(kern, Xa, X2a)
    if X2 is None and kern.get("white") is not None:
        K = K + kern["white"] * torch.eye(X.shape[0], dtype=X.dtype)
    return K


def kern_Kdiag(kern, X):
    if kern["kind"] == "switched":
        return switched_Kdiag(kern, X)
    kd = kern["variance"] * torch.ones(X.shape[0], dtype=X.dtype)
    if kern.get("white") is not None:
        kd = kd + kern["white"]
    return kd


def switched_K(kern, X, X2=None):
    """dgplib/specialized_kernels.py:23-60: the last column is an integer task id; the gram matrix is the
    per-task gram where both task ids agree and zero elsewhere (partition -> per-task K -> scatter -> stitch)."""
    symmetric = X2 is None
    if symmetric:
        X2 = X
    ind = X[:, -1].to(torch.int64)       # :28-29
    ind2 = X2[:, -1].to(torch.int64)     # :31-32
    Xv, X2v = X[:, :-1], X2[:, :-1]      # :34-35
    out = torch.zeros(X.shape[0], X2.shape[0], dtype=X.dtype)
    for t, k in enumerate(kern["kernels"]):
        p = torch.nonzero(ind == t).flatten()     # dynamic_partition, :40-43
        p2 = torch.nonzero(ind2 == t).flatten()
        if p.numel() == 0 or p2.numel() == 0:
            continue
        # :47 calls k.K(gather(X,p), gather(X2,p2)) with an explicit X2, so White adds nothing here.
        gram = kern_K(k, Xv[p], X2v[p2])
        rows = p[:, None].expand(-1, p2.numel())
        cols = p2[None, :].expand(p.numel(), -1)
        out = out.index_put((rows, cols), gram)   # scatter_nd + dynamic_stitch, :50-60
    return out


def switched_Kdiag(kern, X):
    """dgplib/specialized_kernels.py:62-72."""
    ind = X[:, -1].to(torch.int64)
    Xv = X[:, :-1]
    out = torch.zeros(X.shape[0], dtype=X.dtype)
    for t, k in enumerate(kern["kernels"]):
        p = torch.nonzero(ind == t).flatten()
        if p.numel():
            out = out.index_put((p,), kern_Kdiag(k, Xv[p]))
    return out


# ----------------------------------------------------------------------------------- GPflow conditional / KL
def conditional_white(X, Z, kern, q_mu, q_sqrt):
    """gpflow.conditionals.conditional(X, InducingPoints(Z), kern, q_mu, q_sqrt=q_sqrt, full_cov=False, white=True)
    -> base_conditional, as called at dgplib/layers.py:66-69.  Returns mean [N,R], var [N,R]."""
    M = Z.shape[0]
    Kmm = kern_K(kern, Z) + JITTER * torch.eye(M, dtype=Z.dtype)   # features.Kuu(jitter)
    Kmn = kern_K(kern, Z, X)                                        # features.Kuf
    Knn = kern_Kdiag(kern, X)
    Lm = torch.linalg.cholesky(Kmm)
    A = torch.linalg.solve_triangular(Lm, Kmn, upper=False)        # [M,N]
    fvar = Knn - (A * A).sum(0)                                     # [N]
    fmean = A.t() @ q_mu                                            # [N,R]
    L = torch.tril(q_sqrt)                                          # band_part(q_sqrt,-1,0)  [R,M,M]
    R = q_sqrt.shape[0]
    LTA = L.transpose(1, 2) @ A.unsqueeze(0).expand(R, -1, -1)      # [R,M,N] materialised (dense matmul)
    fvar = fvar[None, :] + (LTA * LTA).sum(1)                       # [R,N]
    return fmean, fvar.t()


def conditional_white_full_cov(X, Z, kern, q_mu, q_sqrt):
    """base_conditional(full_cov=True): fvar [R, N, N] = Knn - A^T A + LTA^T LTA with Knn = kern.K(X) (symmetric call,
    so White sits on its diagonal).  Called from dgplib/layers.py:66-69 with full_cov=True."""
    M = Z.shape[0]
    Kmm = kern_K(kern, Z) + JITTER * torch.eye(M, dtype=Z.dtype)
    Kmn = kern_K(kern, Z, X)
    Knn = kern_K(kern, X)
    Lm = torch.linalg.cholesky(Kmm)
    A = torch.linalg.solve_triangular(Lm, Kmn, upper=False)
    fvar = Knn - A.t() @ A
    fmean = A.t() @ q_mu
    L = torch.tril(q_sqrt)
    R = q_sqrt.shape[0]
    LTA = L.transpose(1, 2) @ A.unsqueeze(0).expand(R, -1, -1)
    fvar = fvar[None] + LTA.transpose(1, 2) @ LTA
    return fmean, fvar                                             # [N,R], [R,N,N]


def gauss_kl_white(q_mu, q_sqrt):
    """gpflow.kullback_leiblers.gauss_kl(q_mu, q_sqrt, K=None), called at dgplib/layers.py:60."""
    M, R = q_mu.shape
    Lq = torch.tril(q_sqrt)
    diag = torch.diagonal(Lq, dim1=1, dim2=2)
    return 0.5 * ((q_mu * q_mu).sum() - M * R - torch.log(diag * diag).sum() + (Lq * Lq).sum())


def mean_function(mean, X, R):
    if mean is None:  # gpflow Zero
        return torch.zeros(X.shape[0], R, dtype=X.dtype)
    return X @ mean["A"] + mean["b"]  # gpflow Linear


# ---------------------------------------------------------------------------------------------- dgplib layers
def layer_conditional(layer, X_flat):
    """f_conditional of dgplib/layers.py:65-71 (one group) and dgplib/multikernel_layers.py:66-90 (K groups)."""
    groups = layer["groups"]
    R = layer["q_mu"].shape[1]
    K = len(groups)
    off = R // K                                                   # multikernel_layers.py:40
    means, vars_ = [], []
    for i, g in enumerate(groups):
        m, v = conditional_white(X_flat, g["Z"], g["kernel"],
                                 layer["q_mu"][:, i * off:(i + 1) * off],
                                 layer["q_sqrt"][i * off:(i + 1) * off])
        means.append(m)
        vars_.append(v)
    mean = torch.cat(means, -1)
    var = torch.cat(vars_, -1)
    return mean + mean_function(layer.get("mean"), X_flat, R), var


def layer_build_predict(layer, Xnew):
    """multisample_conditional, diag branch (dgplib/layers.py:82-87): [S,N,D] -> flatten -> conditional -> [S,N,R]."""
    S, N, D = Xnew.shape
    mean, var = layer_conditional(layer, Xnew.reshape(S * N, D))
    return mean.reshape(S, N, -1), var.reshape(S, N, -1)


def layer_build_predict_full_cov(layer, Xnew):
    """multisample_conditional, full_cov branch (dgplib/layers.py:74-81; multikernel_layers.py:79-83, :93-97): map over
    the S samples, each conditional's [R,N,N] covariance transposed to [N,N,R].  Returns mean [S,N,R], var [S,N,N,R]."""
    groups = layer["groups"]
    R = layer["q_mu"].shape[1]
    off = R // len(groups)
    means, vars_ = [], []
    for s in range(Xnew.shape[0]):
        ms, vs = [], []
        for i, g in enumerate(groups):
            m, v = conditional_white_full_cov(Xnew[s], g["Z"], g["kernel"], layer["q_mu"][:, i * off:(i + 1) * off],
                                              layer["q_sqrt"][i * off:(i + 1) * off])
            ms.append(m)
            vs.append(v.permute(2, 1, 0))                          # tf.transpose(v): [R,N,N] -> [N,N,R]
        means.append(torch.cat(ms, -1) + mean_function(layer.get("mean"), Xnew[s], R))
        vars_.append(torch.cat(vs, -1))
    return torch.stack(means), torch.stack(vars_)


def normal_sample_full_cov(mean, var, z):
    """dgplib/utils.py:28-36: mean [S,N,D], var [S,N,N,D], z [S,D,N] -> mean + chol(var + jitter I) z per (s, d)."""
    S, N, D = mean.shape
    v = var.permute(0, 3, 1, 2) + JITTER * torch.eye(N, dtype=mean.dtype)[None, None]
    chol = torch.linalg.cholesky(v)                                 # [S,D,N,N]
    f = mean.permute(0, 2, 1) + (chol @ z.unsqueeze(-1))[..., 0]    # [S,D,N]
    return f.permute(0, 2, 1)


def propagate_full_cov(layers, Xnew, S, z, multitask=False):
    """DSDGP._propagate with full_cov=True (dgplib/dsdgp.py:84-99); z[l] is the N(0,1) draw [S, R_l, N]."""
    Fs = [tile_over_samples(Xnew, S)]
    Fmeans, Fvars = [], []
    for l, layer in enumerate(layers):
        mean, var = layer_build_predict_full_cov(layer, Fs[-1])
        F = normal_sample_full_cov(mean, var, z[l])
        if multitask:
            F = torch.cat([F, Fs[-1][:, :, -1:]], -1)
        Fs.append(F)
        Fmeans.append(mean)
        Fvars.append(var)
    return Fs[1:], Fmeans, Fvars


def predict_f_samples(layers, Xnew, num_samples, z_layers, V, multitask=False):
    """DSDGP.predict_f_samples (dgplib/dsdgp.py:170-185): one full-covariance propagation, then per output i
    mu[:, i] + chol(var[:, :, i] + jitter I) V_i with V [D_Y, N, num_samples]; returns [num_samples, N, D_Y]."""
    _, Fmeans, Fvars = propagate_full_cov(layers, Xnew, 1, z_layers, multitask)
    mu, var = Fmeans[-1][0], Fvars[-1][0]
    N = mu.shape[0]
    out = []
    for i in range(mu.shape[1]):
        L = torch.linalg.cholesky(var[:, :, i] + JITTER * torch.eye(N, dtype=mu.dtype))
        out.append(mu[:, i:i + 1] + L @ V[i])
    return torch.stack(out).permute(2, 1, 0)


def layer_kl(layer):
    return gauss_kl_white(layer["q_mu"], layer["q_sqrt"])         # layers.py:58-60 / multikernel_layers.py:62-63


# ------------------------------------------------------------------------------------------------ dgplib model
def tile_over_samples(X, S):
    return X.unsqueeze(0).repeat(S, *([1] * X.dim()))              # dgplib/utils.py:15-18 (real copies)


def propagate(layers, Xnew, S, eps, multitask=False):
    """DSDGP._propagate (dgplib/dsdgp.py:84-99) / MultitaskDSDGP._propagate (dgplib/multitask_dsdgp.py:9-26).
    eps[l] is the N(0,1) draw [S,N,R_l] that the reference takes unseeded at dgplib/utils.py:26."""
    Fs = [tile_over_samples(Xnew, S)]
    Fmeans, Fvars = [], []
    for l, layer in enumerate(layers):
        mean, var = layer_build_predict(layer, Fs[-1])
        F = mean + eps[l] * var ** 0.5                             # utils.py:27 (no clamp)
        if multitask:
            F = torch.cat([F, Fs[-1][:, :, -1:]], -1)              # multitask_dsdgp.py:20
        Fs.append(F)
        Fmeans.append(mean)
        Fvars.append(var)
    return Fs[1:], Fmeans, Fvars


def gaussian_ve(Fmu, Fvar, Y, variance):
    return -0.5 * math.log(2 * math.pi) - 0.5 * torch.log(variance) - 0.5 * ((Y - Fmu) ** 2 + Fvar) / variance


def variational_expectations(lik, Fmu, Fvar, Y):
    if lik["kind"] == "gaussian":
        return gaussian_ve(Fmu, Fvar, Y, lik["variance"])
    # gpflow SwitchedLikelihood: the last column of Y selects the likelihood row by row.
    ind = Y[:, -1].to(torch.int64)
    Yv = Y[:, :-1]
    out = torch.zeros_like(Fmu)
    for t, v in enumerate(lik["variances"]):
        p = torch.nonzero(ind == t).flatten()
        if p.numel():
            out = out.index_put((p,), gaussian_ve(Fmu[p], Fvar[p], Yv[p], v))
    return out


def elbo(layers, lik, X, Y, S, eps, num_data=None, multitask=False):
    """DSDGP._build_likelihood (dgplib/dsdgp.py:107-130)."""
    _, Fmeans, Fvars = propagate(layers, X, S, eps, multitask)
    Fmean, Fvar = Fmeans[-1], Fvars[-1]
    Yt = tile_over_samples(Y, S)
    var_exp = torch.stack([variational_expectations(lik, Fmean[s], Fvar[s], Yt[s]) for s in range(S)])  # map_fn
    L = var_exp.mean(0).sum()                                       # :121-122
    KL = sum(layer_kl(layer) for layer in layers)                   # :124-126
    num_data = X.shape[0] if num_data is None else num_data
    scale = float(num_data) / float(X.shape[0])                     # :128-129
    return L * scale - KL


def predict_f(layers, Xnew, S, eps, multitask=False):
    """DSDGP.predict_f (dgplib/dsdgp.py:133-139) -> last layer's (mean, var), each [S, N*, D_Y]."""
    _, Fmeans, Fvars = propagate(layers, Xnew, S, eps, multitask)
    return Fmeans[-1], Fvars[-1]


def predict_all_layers(layers, Xnew, S, eps, multitask=False):
    return propagate(layers, Xnew, S, eps, multitask)              # dgplib/dsdgp.py:152-158


def predict_y(layers, lik, Xnew, eps, S=1, multitask=False):
    """DSDGP.predict_y (dgplib/dsdgp.py:188-194): single-sample propagate, then likelihood.predict_mean_and_var."""
    mean, var = predict_f(layers, Xnew, S, eps, multitask)
    if lik["kind"] == "gaussian":
        return mean, var + lik["variance"]
    # gpflow SwitchedLikelihood.predict_mean_and_var concatenates every likelihood's prediction along axis 1
    # (pinned by the reference's tests/test_multitask_dsdgp.py:165-173: 10 test rows -> shape (1, 20, 1)).
    mus = [mean for _ in lik["variances"]]
    vs = [var + v for v in lik["variances"]]
    return torch.cat(mus, 1), torch.cat(vs, 1)


# -------------------------------------------------------------------------------------------- host-side init
def find_weights(input_dim, output_dim, X, multitask=False):
    """dgplib/layers.py:97-121: initial weights of a Linear mean function (identity / top right-singular vectors /
    zero padding); multitask appends an all-zero row for the task-id column."""
    import numpy as np
    if input_dim == output_dim:
        W = np.eye(input_dim)
    elif input_dim > output_dim:
        Xv = X[:, :-1] if multitask else X
        W = np.linalg.svd(Xv, full_matrices=False)[2][:output_dim].T
    else:
        W = np.hstack([np.eye(input_dim), np.zeros((input_dim, output_dim - input_dim))])
    if multitask:
        W = np.vstack([W, np.zeros((1, W.shape[1]))])
    return W


# --------------------------------------------------------------------------------------------- autograd helper
def leaves(layers, lik):
    """All gradient targets: per layer/group Z, kernel variance / lengthscales / white; q_mu, q_sqrt;
    likelihood variance(s).  Returns an ordered dict name -> tensor."""
    out = {}

    def kern_leaves(prefix, k):
        if k["kind"] == "switched":
            for t, kk in enumerate(k["kernels"]):
                kern_leaves(f"{prefix}.k{t}", kk)
            return
        out[f"{prefix}.variance"] = k["variance"]
        out[f"{prefix}.lengthscales"] = k["lengthscales"]
        if k.get("white") is not None:
            out[f"{prefix}.white"] = k["white"]

    for l, layer in enumerate(layers):
        seen = {}
        for g, grp in enumerate(layer["groups"]):
            if id(grp["Z"]) not in seen:
                seen[id(grp["Z"])] = g
                out[f"l{l}.g{g}.Z"] = grp["Z"]
            kern_leaves(f"l{l}.g{g}.kern", grp["kernel"])
        out[f"l{l}.q_mu"] = layer["q_mu"]
        out[f"l{l}.q_sqrt"] = layer["q_sqrt"]
        if layer.get("mean") is not None and layer["mean"].get("trainable"):
            # only an OutputLayer built with its own Linear mean keeps it trainable (dgplib/layers.py:211-218 does not
            # freeze it, unlike :184-186 and :202-204)
            out[f"l{l}.mean.A"] = layer["mean"]["A"]
            out[f"l{l}.mean.b"] = layer["mean"]["b"]
    if lik["kind"] == "gaussian":
        out["lik.variance"] = lik["variance"]
    else:
        for t, v in enumerate(lik["variances"]):
            out[f"lik.variance{t}"] = v
    return out


def elbo_and_grad(layers, lik, X, Y, S, eps, num_data=None, multitask=False):
    lv = leaves(layers, lik)
    for t in lv.values():
        t.requires_grad_(True)
        t.grad = None
    val = elbo(layers, lik, X, Y, S, eps, num_data, multitask)
    grads = torch.autograd.grad(val, list(lv.values()), allow_unused=True)
    g = {k: (torch.zeros_like(t) if gr is None else gr) for (k, t), gr in zip(lv.items(), grads)}
    for t in lv.values():
        t.requires_grad_(False)
    return val.detach(), g
