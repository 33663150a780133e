"""50-digit arbiter for BASELINE configs[0] with the notebooks' White(1e-5): ELBO and its gradient w.r.t. every kernel
hyper-parameter and the likelihood variance of the 2-layer RBF model of tests/test_gpu_model.py::test_c1_jitter_only_
conditioning (M = 50, D = 1, N = 1000, S = 1; cond(Kuu) ~ 4e6).  TEST INFRASTRUCTURE ONLY.

Why: at this conditioning the torch oracle and the CUDA path both carry ~cond * eps of rounding in the White-variance
gradient and differ by ~1e-9 of each other; this script decides which side is closer to the exact value.  The gradient
is a central difference of the 50-digit ELBO (oracle/mp_arbiter.py's formulas, restated with dot products for speed)
with relative step 1e-18: truncation ~1e-36, rounding ~1e-32 -- exact for the purposes of a 1e-9 comparison.

    python -m oracle.mp_c1_grad            # ~10 min on one core; writes tests/golden/c1_white_grad_mp.json
"""
import json
import os
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 50
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cond(F, layer):
    """Whitened conditional of an RBF + White layer at the rows F (list of lists of mpf); SURVEY.md Appendix A."""
    Z, var, ls, white = layer["Z"], layer["variance"], layer["lengthscale"], layer["white"]
    M, N, D = len(Z), len(F), len(Z[0])
    zs = [[z / ls for z in row] for row in Z]
    zn = [mp.fdot(r, r) for r in zs]
    K = mp.matrix(M, M)
    for i in range(M):
        for j in range(i + 1):
            r2 = zn[i] + zn[j] - 2 * mp.fdot(zs[i], zs[j])
            K[i, j] = K[j, i] = var * mp.exp(-r2 / 2)
        K[i, i] += white + mp.mpf("1e-6")
    Lm = mp.cholesky(K)
    Lrows = [[Lm[i, j] for j in range(i)] for i in range(M)]
    Ldiag = [Lm[i, i] for i in range(M)]
    q_mu, Lq = layer["q_mu"], layer["q_sqrt"]
    R = len(q_mu[0])
    qcols = [[q_mu[m][r] for m in range(M)] for r in range(R)]
    # columns of tril(q_sqrt_r): LqT[r][m] = [Lq[r][k][m] for k >= m]
    LqT = [[[Lq[r][k][m] for k in range(m, M)] for m in range(M)] for r in range(R)]
    mean, fvar = [], []
    for n in range(N):
        xs = [x / ls for x in F[n]]
        xn = mp.fdot(xs, xs)
        kcol = [var * mp.exp(-(zn[m] + xn - 2 * mp.fdot(zs[m], xs)) / 2) for m in range(M)]
        a = []
        for m in range(M):                      # forward substitution Lm a = k
            a.append((kcol[m] - mp.fdot(Lrows[m], a)) / Ldiag[m])
        base = var + white - mp.fdot(a, a)
        mrow, vrow = [], []
        for r in range(R):
            mu = mp.fdot(a, qcols[r])
            if layer.get("mean_A") is not None:
                mu += mp.fdot(F[n], [layer["mean_A"][k][r] for k in range(len(F[n]))])
            acc = mp.mpf(0)
            for m in range(M):
                u = mp.fdot(LqT[r][m], a[m:])
                acc += u * u
            mrow.append(mu)
            vrow.append(base + acc)
        mean.append(mrow)
        fvar.append(vrow)
    return mean, fvar


def _kl(layer):
    q_mu, Lq = layer["q_mu"], layer["q_sqrt"]
    M, R = len(q_mu), len(q_mu[0])
    s = sum(q_mu[m][r] ** 2 for m in range(M) for r in range(R)) - M * R
    for r in range(R):
        for i in range(M):
            s += mp.fdot(Lq[r][i][:i + 1], Lq[r][i][:i + 1]) - mp.log(Lq[r][i][i] ** 2)
    return s / 2


def elbo(model):
    X, Y, eps = model["X"], model["Y"], model["eps"]
    N = len(X)
    lik = model["lik_variance"]
    F = X
    for l, layer in enumerate(model["layers"]):
        mean, var = _cond(F, layer)
        F = [[mean[n][r] + eps[l][n][r] * mp.sqrt(var[n][r]) for r in range(len(mean[n]))] for n in range(N)]
    total = mp.mpf(0)
    for n in range(N):
        for r in range(len(mean[n])):
            total += -mp.log(2 * mp.pi) / 2 - mp.log(lik) / 2 - ((Y[n][r] - mean[n][r]) ** 2 + var[n][r]) / (2 * lik)
    return total * model["num_data"] / N - sum(_kl(layer) for layer in model["layers"])


def _mpf_tree(x):
    if isinstance(x, (list, tuple)):
        return [_mpf_tree(v) for v in x]
    return mp.mpf(float(x))


def c1_model(N=1000, white=1e-5):
    """Parameter VALUES of tests/_models.build_dsdgp(L=2, M=50, D=1, N=N, S=1, white=white) as exact binary doubles."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, ROOT)
    from _models import build_dsdgp, make_eps
    model, X, Y = build_dsdgp(L=2, M=50, D=1, N=N, S=1, white=white)
    eps = make_eps(model, 1, N)
    layers = []
    for layer in model.layers.layers:
        (k, f, c0, R), = layer.conditionals()
        (kind, var, wh, ls), = k.task_kernels()
        mf = layer.mean_function
        layers.append({"Z": f.Z.value.tolist(), "variance": float(var), "lengthscale": float(ls.reshape(-1)[0]),
                       "white": float(wh), "q_mu": layer.q_mu.value.tolist(),
                       "q_sqrt": np.tril(layer.q_sqrt.value).tolist(),
                       "mean_A": mf.A.value.tolist() if type(mf).__name__ == "Linear" else None})
    return {"X": X.tolist(), "Y": Y.tolist(), "eps": [e[0].tolist() for e in eps], "num_data": N,
            "lik_variance": float(model.likelihood.variance.value), "layers": layers}


def to_mp(spec):
    m = {"X": _mpf_tree(spec["X"]), "Y": _mpf_tree(spec["Y"]), "eps": _mpf_tree(spec["eps"]),
         "num_data": mp.mpf(spec["num_data"]), "lik_variance": mp.mpf(spec["lik_variance"]), "layers": []}
    for L in spec["layers"]:
        m["layers"].append({"Z": _mpf_tree(L["Z"]), "variance": mp.mpf(L["variance"]),
                            "lengthscale": mp.mpf(L["lengthscale"]), "white": mp.mpf(L["white"]),
                            "q_mu": _mpf_tree(L["q_mu"]), "q_sqrt": _mpf_tree(L["q_sqrt"]),
                            "mean_A": _mpf_tree(L["mean_A"]) if L["mean_A"] is not None else None})
    return m


def gradient(m, names):
    """Central differences of the 50-digit ELBO; names: list of (layer index or None, key)."""
    out = {}
    h_rel = mp.mpf("1e-18")
    for (l, key) in names:
        tgt = m if l is None else m["layers"][l]
        v0 = tgt[key]
        h = v0 * h_rel
        tgt[key] = v0 + h
        fp = elbo(m)
        tgt[key] = v0 - h
        fm = elbo(m)
        tgt[key] = v0
        out[(f"l{l}." if l is not None else "") + key] = (fp - fm) / (2 * h)
        print((l, key), mp.nstr(out[(f"l{l}." if l is not None else "") + key], 25), flush=True)
    return out


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    spec = c1_model(N)
    m = to_mp(spec)
    val = elbo(m)
    print("elbo", mp.nstr(val, 30), flush=True)
    names = [(l, k) for l in range(2) for k in ("white", "variance", "lengthscale")] + [(None, "lik_variance")]
    g = gradient(m, names)
    out = os.path.join(ROOT, "tests", "golden", "c1_white_grad_mp.json")
    with open(out, "w") as f:
        json.dump({"generator": f"python -m oracle.mp_c1_grad {N}", "digits": 50,
                   "model": "tests/_models.build_dsdgp(L=2, M=50, D=1, N=%d, S=1, white=1e-5), eps = make_eps(model, 1, N)" % N,
                   "N": N, "elbo_mp": mp.nstr(val, 30), "grad_mp": {k: mp.nstr(v, 30) for k, v in g.items()}}, f, indent=1)
    print("wrote", out)
