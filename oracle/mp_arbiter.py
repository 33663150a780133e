"""50-digit mpmath evaluation of the DSDGP ELBO on tiny models: the arbiter at the 1e-9 tolerance and the source of the
high-precision fixture tests/golden/tiny_elbo_mp.json.  TEST INFRASTRUCTURE ONLY.

Independent of the torch oracle: plain Python loops over mpmath numbers, following SURVEY.md Appendix A
(RBF with White, whitened conditional, whitened KL, Gaussian variational expectation, dgplib/dsdgp.py:107-130)."""
import json
import os

import mpmath as mp
import numpy as np

mp.mp.dps = 50


def _rbf(X, X2, var, ls):
    out = mp.matrix(len(X), len(X2))
    for i, x in enumerate(X):
        for j, z in enumerate(X2):
            xs = [mp.mpf(a) / ls for a in x]
            zs = [mp.mpf(a) / ls for a in z]
            r2 = -2 * sum(a * b for a, b in zip(xs, zs)) + sum(a * a for a in xs) + sum(b * b for b in zs)
            out[i, j] = var * mp.exp(-r2 / 2)
    return out


def _conditional(X, layer):
    Z, var, ls, white = layer["Z"], mp.mpf(layer["variance"]), mp.mpf(layer["lengthscale"]), mp.mpf(layer["white"])
    M, N = len(Z), len(X)
    Kmm = _rbf(Z, Z, var, ls)
    for i in range(M):
        Kmm[i, i] += white + mp.mpf("1e-6")
    Lm = mp.cholesky(Kmm)
    Kmn = _rbf(Z, X, var, ls)
    A = mp.inverse(Lm) * Kmn
    q_mu, q_sqrt = layer["q_mu"], layer["q_sqrt"]
    R = len(q_mu[0])
    mean = [[sum(A[m, n] * mp.mpf(q_mu[m][r]) for m in range(M)) for r in range(R)] for n in range(N)]
    fvar = []
    for n in range(N):
        base = var + white - sum(A[m, n] ** 2 for m in range(M))
        row = []
        for r in range(R):
            acc = mp.mpf(0)
            for m in range(M):   # (Lq^T A)[m, n] = sum_{k >= m} Lq[k, m] A[k, n]
                u = sum(mp.mpf(q_sqrt[r][k][m]) * A[k, n] for k in range(m, M))
                acc += u * u
            row.append(base + acc)
        fvar.append(row)
    if layer.get("mean_A") is not None:
        for n in range(N):
            for r in range(R):
                mean[n][r] += sum(mp.mpf(X[n][k]) * mp.mpf(layer["mean_A"][k][r]) for k in range(len(X[n])))
    return mean, fvar


def _kl(layer):
    q_mu, q_sqrt = layer["q_mu"], layer["q_sqrt"]
    M, R = len(q_mu), len(q_mu[0])
    s = sum(mp.mpf(q_mu[m][r]) ** 2 for m in range(M) for r in range(R)) - M * R
    for r in range(R):
        for i in range(M):
            for j in range(i + 1):
                s += mp.mpf(q_sqrt[r][i][j]) ** 2
            s -= mp.log(mp.mpf(q_sqrt[r][i][i]) ** 2)
    return s / 2


def elbo(model):
    X, Y, S, eps = model["X"], model["Y"], model["S"], model["eps"]
    N = len(X)
    lik = mp.mpf(model["lik_variance"])
    total = mp.mpf(0)
    for s in range(S):
        F = [[mp.mpf(v) for v in row] for row in X]
        for l, layer in enumerate(model["layers"]):
            mean, var = _conditional(F, layer)
            F = [[mean[n][r] + mp.mpf(eps[l][s][n][r]) * mp.sqrt(var[n][r]) for r in range(len(mean[n]))]
                 for n in range(N)]
        for n in range(N):
            for r in range(len(mean[n])):
                total += (-mp.log(2 * mp.pi) / 2 - mp.log(lik) / 2
                          - ((mp.mpf(Y[n][r]) - mean[n][r]) ** 2 + var[n][r]) / (2 * lik)) / S
    scale = mp.mpf(model["num_data"]) / N
    return total * scale - sum(_kl(layer) for layer in model["layers"])


def tiny_model(seed=0, M=4, N=5, D=2, S=2):
    rng = np.random.RandomState(seed)
    X = rng.randn(N, D)
    layers = []
    dims = [D, D, 1]
    for l in range(2):
        R = dims[l + 1]
        layers.append({"Z": X[:M].tolist() if l == 0 else rng.randn(M, D).tolist(),
                       "variance": 1.3 - 0.2 * l, "lengthscale": 1.1 + 0.3 * l, "white": 1e-3,
                       "q_mu": (0.3 * rng.randn(M, R)).tolist(),
                       "q_sqrt": (np.eye(M)[None] + 0.2 * np.tril(rng.randn(R, M, M))).tolist(),
                       "mean_A": np.eye(D).tolist() if l == 0 else None})
    return {"X": X.tolist(), "Y": rng.randn(N, 1).tolist(), "S": S, "num_data": 3 * N, "lik_variance": 0.2,
            "layers": layers,
            "eps": [rng.randn(S, N, dims[l + 1]).tolist() for l in range(2)]}


if __name__ == "__main__":
    # regenerates the committed fixture
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tiny_elbo_mp.json")
    models = []
    for seed in (0, 1):
        m = tiny_model(seed)
        m["elbo_mp"] = mp.nstr(elbo(m), 30)
        models.append(m)
    with open(out, "w") as f:
        json.dump({"generator": "python -m oracle.mp_arbiter", "digits": 50, "models": models}, f)
    print("wrote", out, [m["elbo_mp"] for m in models])
