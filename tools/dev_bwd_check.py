"""Dev check (GPU): raw C-ABI backward of one conditional vs torch autograd of the oracle."""
import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgplib_b200 import _cabi as C
from oracle import dsdgp_oracle as O

def run(M, D, R, N, kind="rbf", white=1e-5, seed=0, linear=True, use_gF=True):
    rng = np.random.RandomState(seed)
    X = torch.tensor(rng.randn(N, D)); Z = torch.tensor(rng.randn(M, D))
    ls = torch.tensor(np.sqrt(D) * (0.8 + 0.4 * rng.rand(D)))
    var = torch.tensor(1.3, dtype=torch.float64); wh = torch.tensor(white, dtype=torch.float64)
    q_mu = torch.tensor(0.1 * rng.randn(M, R)); q_sqrt = torch.tensor(np.eye(M)[None] + 0.1 * np.tril(rng.randn(R, M, M)))
    A = torch.tensor(rng.randn(D, R)) if linear else None; b = torch.tensor(rng.randn(R)) if linear else None
    eps = torch.tensor(rng.randn(N, R))
    gm_in = torch.tensor(rng.randn(N, R)); gv_in = torch.tensor(rng.randn(N, R)); gF = torch.tensor(rng.randn(N, R))
    kern = {"kind": kind, "input_dim": D, "variance": var, "lengthscales": ls, "white": wh}
    layer = {"groups": [{"kernel": kern, "Z": Z}], "q_mu": q_mu, "q_sqrt": q_sqrt, "mean": {"A": A, "b": b} if linear else None}
    leaves = [X, Z, var, ls, wh, q_mu, q_sqrt]
    for t in leaves: t.requires_grad_(True)
    m_ref, v_ref = O.layer_build_predict(layer, X[None]); m_ref = m_ref[0]; v_ref = v_ref[0]
    F_ref = m_ref + eps * v_ref ** 0.5
    obj = (gm_in * m_ref).sum() + (gv_in * v_ref).sum() - O.layer_kl(layer)
    if use_gF: obj = obj + (gF * F_ref).sum()
    grads = torch.autograd.grad(obj, leaves)
    for t in leaves: t.requires_grad_(False)
    ref = dict(zip(["X", "Z", "var", "ls", "wh", "q_mu", "q_sqrt"], grads))
    dev = "cuda"; lib = C.lib()
    d = C.CondDesc(); d.M = M; d.D = D; d.Dx = D; d.R = R; d.ldr = R; d.col0 = 0; d.T = 1
    d.kind[0] = 0 if kind == "rbf" else 1; d.jitter = 1e-6
    theta = torch.cat([var.reshape(1), wh.reshape(1), ls]).detach().to(dev)
    Zd, qmd, qsd, Xd, epsd = Z.detach().to(dev), q_mu.detach().to(dev), q_sqrt.detach().to(dev), X.detach().to(dev), eps.to(dev)
    Ad = A.to(dev) if linear else None; bd = b.to(dev) if linear else None
    d.Z = Zd.data_ptr(); d.theta = theta.data_ptr(); d.q_mu = qmd.data_ptr(); d.q_sqrt = qsd.data_ptr()
    d.mean_A = Ad.data_ptr() if linear else None; d.mean_b = bd.data_ptr() if linear else None
    wsb = lib.dgp_cond_ws_bytes(ctypes.byref(d), N, 1)
    ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
    kl = torch.zeros(1, dtype=torch.float64, device=dev)
    C.check(lib.dgp_cond_prepare(ctypes.byref(d), C.ptr(ws), wsb, 1, C.ptr(kl), C.stream_ptr()), "prepare")
    Mp = C.mpad(M)
    mean = torch.zeros(N, R, dtype=torch.float64, device=dev); varo = torch.zeros_like(mean); F = torch.zeros_like(mean)
    Asave = torch.zeros(N, Mp, dtype=torch.float64, device=dev)
    C.check(lib.dgp_cond_forward(ctypes.byref(d), C.ptr(ws), C.ptr(Xd), N, N, C.ptr(epsd), 0, None, 0, 0, 0, C.ptr(mean), C.ptr(varo), C.ptr(F), R, C.ptr(Asave), C.stream_ptr()), "forward")
    gX = torch.zeros(N, D, dtype=torch.float64, device=dev)
    gmd, gvd, gFd = gm_in.to(dev), gv_in.to(dev), gF.to(dev)
    C.check(lib.dgp_cond_backward(ctypes.byref(d), C.ptr(ws), C.ptr(Xd), N, N, C.ptr(Asave), C.ptr(mean), C.ptr(varo), C.ptr(F), R,
                                  C.ptr(gmd), C.ptr(gvd), C.ptr(gFd) if use_gF else None, R, C.ptr(gX), 0, C.stream_ptr()), "backward")
    C.check(lib.dgp_cond_backward_gram(ctypes.byref(d), C.ptr(ws), C.ptr(Asave), N, 0, C.stream_ptr()), "gram")
    gZ = torch.zeros(M, D, dtype=torch.float64, device=dev); gth = torch.zeros(2 + D, dtype=torch.float64, device=dev)
    gqm = torch.zeros(M, R, dtype=torch.float64, device=dev); gqs = torch.zeros(R, M, M, dtype=torch.float64, device=dev)
    C.check(lib.dgp_cond_backward_params(ctypes.byref(d), C.ptr(ws), 1.0, C.ptr(gZ), C.ptr(gth), C.ptr(gqm), C.ptr(gqs), C.stream_ptr()), "params")
    torch.cuda.synchronize()
    rel = lambda a, b: float((a.cpu() - b).abs().max() / max(float(b.abs().max()), 1e-300))
    print(f"M={M} D={D} R={R} N={N} {kind} gF={use_gF}: gX={rel(gX, ref['X']):.2e} gZ={rel(gZ, ref['Z']):.2e} gvar={rel(gth[0:1], ref['var'].reshape(1)):.2e} "
          f"gwh={rel(gth[1:2], ref['wh'].reshape(1)):.2e} gls={rel(gth[2:], ref['ls']):.2e} gq_mu={rel(gqm, ref['q_mu']):.2e} gq_sqrt={rel(gqs, ref['q_sqrt']):.2e}", flush=True)

if __name__ == "__main__":
    run(50, 1, 1, 200, use_gF=False)
    run(64, 3, 2, 130)
    run(100, 8, 8, 777, kind="matern52")
    run(128, 8, 8, 4100)
    run(256, 8, 8, 5000)
    run(300, 16, 4, 1000, kind="matern52", linear=False)
    run(500, 16, 16, 2000, kind="matern52")
