"""Copies parameter VALUES out of a dgplib_b200 model into the plain-dict description the oracle consumes.
TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference)."""
import numpy as np
import torch

T64 = torch.float64


def t(x):
    return torch.tensor(np.array(x, dtype=np.float64), dtype=T64)


def oracle_kernel(k):
    name = type(k).__name__
    if name == "SwitchedKernel":
        return {"kind": "switched", "kernels": [oracle_kernel(kk) for kk in k.kernels]}
    stat, white = k, None
    if name == "Sum":
        stat = [kk for kk in k.kernels if type(kk).__name__ in ("RBF", "Matern52")][0]
        ws = [kk for kk in k.kernels if type(kk).__name__ == "White"]
        white = t(sum(float(w.variance.value) for w in ws)) if ws else None
    return {"kind": "rbf" if type(stat).__name__ == "RBF" else "matern52", "input_dim": stat.input_dim,
            "variance": t(stat.variance.value), "lengthscales": t(stat.lengthscales.value), "white": white}


def oracle_layers(model_layers):
    out = []
    for layer in model_layers:
        groups, zcache = [], {}
        for (k, f, c0, R) in layer.conditionals():
            if id(f) not in zcache:
                zcache[id(f)] = t(f.Z.value)
            groups.append({"kernel": oracle_kernel(k), "Z": zcache[id(f)]})
        mean = None
        if type(layer.mean_function).__name__ == "Linear":
            mf = layer.mean_function
            mean = {"A": t(mf.A.value), "b": t(mf.b.value), "trainable": bool(mf.A.trainable or mf.b.trainable)}
        out.append({"groups": groups, "q_mu": t(layer.q_mu.value), "q_sqrt": t(layer.q_sqrt.value), "mean": mean})
    return out


def oracle_lik(lik):
    if hasattr(lik, "likelihood_list"):
        return {"kind": "switched", "variances": [t(l.variance.value) for l in lik.likelihood_list]}
    return {"kind": "gaussian", "variance": t(lik.variance.value)}
