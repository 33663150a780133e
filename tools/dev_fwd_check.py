"""Dev check (GPU): raw C-ABI prepare + forward vs the oracle for one conditional."""
import ctypes, sys, os, math
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgplib_b200 import _cabi as C
from oracle import dsdgp_oracle as O

def run(M, D, R, N, kind="rbf", S=1, white=1e-5, seed=0, linear=True):
    rng = np.random.RandomState(seed)
    X = torch.tensor(rng.randn(N, D)); Z = torch.tensor(rng.randn(M, D))
    ls = torch.tensor(np.sqrt(D) * (0.8 + 0.4 * rng.rand(D)))
    var = torch.tensor(1.3, dtype=torch.float64); wh = torch.tensor(white, dtype=torch.float64)
    q_mu = torch.tensor(0.1 * rng.randn(M, R)); q_sqrt = torch.tensor(np.eye(M)[None] + 0.1 * np.tril(rng.randn(R, M, M)) + 0.05*np.triu(rng.randn(R,M,M),1))
    A = torch.tensor(rng.randn(D, R)) if linear else None; b = torch.tensor(rng.randn(R)) if linear else None
    eps = torch.tensor(rng.randn(S * N, R))
    kern = {"kind": kind, "input_dim": D, "variance": var, "lengthscales": ls, "white": wh}
    layer = {"groups": [{"kernel": kern, "Z": Z}], "q_mu": q_mu, "q_sqrt": q_sqrt, "mean": {"A": A, "b": b} if linear else None}
    Xt = X[None].repeat(S, 1, 1)
    m_ref, v_ref = O.layer_build_predict(layer, Xt)
    m_ref = m_ref.reshape(S * N, R); v_ref = v_ref.reshape(S * N, R)
    F_ref = m_ref + eps * v_ref ** 0.5
    kl_ref = O.layer_kl(layer)
    dev = "cuda"
    lib = C.lib()
    d = C.CondDesc(); d.M = M; d.D = D; d.Dx = D; d.R = R; d.ldr = R; d.col0 = 0; d.T = 1
    d.kind[0] = 0 if kind == "rbf" else 1; d.jitter = 1e-6
    theta = torch.cat([var.reshape(1), wh.reshape(1), ls]).to(dev)
    Zd, qmd, qsd, Xd, epsd = Z.to(dev), q_mu.to(dev), q_sqrt.to(dev), X.to(dev), eps.to(dev)
    Ad = A.to(dev) if linear else None; bd = b.to(dev) if linear else None
    d.Z = Zd.data_ptr(); d.theta = theta.data_ptr(); d.q_mu = qmd.data_ptr(); d.q_sqrt = qsd.data_ptr()
    d.mean_A = Ad.data_ptr() if linear else None; d.mean_b = bd.data_ptr() if linear else None
    nrows = S * N
    wsb = lib.dgp_cond_ws_bytes(ctypes.byref(d), nrows, 0)
    ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
    kl = torch.zeros(1, dtype=torch.float64, device=dev)
    C.check(lib.dgp_cond_prepare(ctypes.byref(d), C.ptr(ws), wsb, 0, C.ptr(kl), C.stream_ptr()), "prepare")
    st = ctypes.c_int32(-1); C.check(lib.dgp_cond_status(C.ptr(ws), C.stream_ptr(), ctypes.byref(st)), "status")
    Mp = C.mpad(M)
    mean = torch.zeros(nrows, R, dtype=torch.float64, device=dev); varo = torch.zeros_like(mean); F = torch.zeros_like(mean)
    Asave = torch.zeros(nrows, Mp, dtype=torch.float64, device=dev)
    C.check(lib.dgp_cond_forward(ctypes.byref(d), C.ptr(ws), C.ptr(Xd), N, nrows, C.ptr(epsd), 0, None, 0, 0, 0, C.ptr(mean), C.ptr(varo), C.ptr(F), R, C.ptr(Asave), C.stream_ptr()), "forward")
    torch.cuda.synchronize()
    rel = lambda a, b: float((a.cpu() - b).abs().max() / b.abs().max())
    # A reference
    Kmm = O.kern_K(kern, Z) + 1e-6 * torch.eye(M, dtype=torch.float64); Lm = torch.linalg.cholesky(Kmm)
    A_ref = torch.linalg.solve_triangular(Lm, O.kern_K(kern, Z, X), upper=False).t()  # [N, M]
    print(f"M={M} D={D} R={R} N={N} S={S} {kind}: status={st.value} kl rel={abs(kl.item()-kl_ref.item())/max(abs(kl_ref.item()),1e-300):.2e} "
          f"A={rel(Asave[:N,:M], A_ref):.2e} padA={float(Asave[:, M:].abs().max()) if Mp>M else 0:.1e} mean={rel(mean, m_ref):.2e} var={rel(varo, v_ref):.2e} F={rel(F, F_ref):.2e}", flush=True)

if __name__ == "__main__":
    if os.environ.get("DGP_DEV_NOBWD"):
        for k in ("dgp_cond_backward", "dgp_cond_backward_params"): C.SIGNATURES.pop(k)
    print(C.lib().dgp_version().decode())
    run(50, 1, 1, 1000)
    run(64, 3, 2, 130, S=2)
    run(100, 8, 8, 777, kind="matern52")
    run(128, 8, 8, 4096, S=2)
    run(256, 8, 8, 5000)
    run(300, 16, 4, 1000, kind="matern52")
    run(500, 16, 16, 2000, kind="matern52")
    run(1000, 8, 2, 300, linear=False)
