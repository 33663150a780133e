"""Dev check (GPU): full-model ELBO + gradients vs the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from _models import *
from oracle import dsdgp_oracle as O

def check(**kw):
    model, X, Y = build_dsdgp(**kw)
    S = model.num_samples
    eps = make_eps(model, S, X.shape[0])
    model.compile()
    val, grads = model.elbo_and_grad(eps=eps)
    ol, olik = oracle_layers(model.layers.layers), oracle_lik(model.likelihood)
    oval, og = O.elbo_and_grad(ol, olik, t(X), t(Y), S, [t(e) for e in eps], num_data=X.shape[0])
    gp = oracle_grads_by_param(model, grads)
    errs = {"elbo": abs(float(val) - float(oval)) / abs(float(oval))}
    for l, layer in enumerate(ol):
        errs[f"l{l}.Z"] = rel_err(gp[f"l{l}.g0.Z"], og[f"l{l}.g0.Z"])
        errs[f"l{l}.q_mu"] = rel_err(gp[f"l{l}.q_mu"], og[f"l{l}.q_mu"])
        errs[f"l{l}.q_sqrt"] = rel_err(gp[f"l{l}.q_sqrt"], og[f"l{l}.q_sqrt"])
        th = gp[f"l{l}.g0.theta"][0].cpu()
        errs[f"l{l}.var"] = rel_err(th[0], og[f"l{l}.g0.kern.variance"])
        errs[f"l{l}.white"] = rel_err(th[1], og[f"l{l}.g0.kern.white"])
        ols = og[f"l{l}.g0.kern.lengthscales"]
        errs[f"l{l}.ls"] = rel_err(th[2:].sum() if ols.dim() == 0 else th[2:], ols)
    errs["lik"] = rel_err(gp["lik"][0], og["lik.variance"])
    worst = max(errs, key=errs.get)
    print(kw, f"elbo={float(val):.6f} worst {worst}={errs[worst]:.2e}", flush=True)
    return errs

if __name__ == "__main__":
    check(L=1, M=20, D=2, N=50, S=1)
    check(L=2, M=50, D=1, N=1000, S=1)
    check(L=3, M=32, D=4, N=200, S=3)
    check(L=3, M=128, D=8, N=500, S=4, kind="matern52", ard=True)
    check(L=3, M=64, D=5, N=300, S=2, hidden=3)
    # autograd through torch
    model, X, Y = build_dsdgp(L=2, M=16, D=2, N=64, S=2)
    model.compile()
    eps = make_eps(model, 2, 64)
    loss = -model._build_likelihood(eps=eps)
    loss.backward()
    n = sum(p.grad is not None for p in model.parameters() if p.requires_grad)
    print("autograd ok, params with grad:", n, "of", sum(p.requires_grad for p in model.parameters()))
