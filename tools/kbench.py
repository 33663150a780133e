"""Kernel micro-benchmark (GPU): times dgp_cond_forward / backward / gram for one conditional shape, CUDA events, best
and mean of several repetitions.  Used for A/B comparisons of kernel variants (DGP_B200_LIB selects the build)."""
import argparse, ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgplib_b200 import _cabi as C
if os.environ.get("DGP_B200_LIB"):          # A/B builds (make prof_lib / test_lib): a tool-side switch, not the product's
    C.use_library(os.environ["DGP_B200_LIB"])

ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=256); ap.add_argument("--D", type=int, default=8)
ap.add_argument("--R", type=int, default=8); ap.add_argument("--N", type=int, default=100000)
ap.add_argument("--reps", type=int, default=8); ap.add_argument("--kind", default="rbf")
ap.add_argument("--debug", type=int, default=0); ap.add_argument("--stages", type=int, default=0); ap.add_argument("--force-fwd", type=int, default=0); ap.add_argument("--force-bwd", type=int, default=0)   # test-hook library only
a = ap.parse_args()
M, D, R, N = a.M, a.D, a.R, a.N
dev = "cuda"; lib = C.lib(); rng = np.random.RandomState(0)
if a.force_fwd or a.force_bwd:
    lib.dgp_test_force_tiles(a.force_fwd, a.force_bwd)
if a.debug:
    lib.dgp_test_debug_mode(a.debug)
if a.stages:
    lib.dgp_test_force_stages(a.stages)
t = lambda x: torch.tensor(x, dtype=torch.float64, device=dev)
X = t(rng.randn(N, D)); Z = t(rng.randn(M, D)); theta = t(np.concatenate([[1.0, 1e-5], np.sqrt(D) * np.ones(D)]))
q_mu = t(0.1 * rng.randn(M, R)); q_sqrt = t(np.eye(M)[None] + 0.1 * np.tril(rng.randn(R, M, M)))
d = C.CondDesc(); d.M, d.D, d.Dx, d.R, d.ldr, d.col0, d.T = M, D, D, R, R, 0, 1
d.kind[0] = 0 if a.kind == "rbf" else 1; d.jitter = 1e-6
d.Z, d.theta, d.q_mu, d.q_sqrt = Z.data_ptr(), theta.data_ptr(), q_mu.data_ptr(), q_sqrt.data_ptr()
wsb = lib.dgp_cond_ws_bytes(ctypes.byref(d), N, 1); ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
st = C.stream_ptr()
C.check(lib.dgp_cond_prepare(ctypes.byref(d), C.ptr(ws), wsb, 1, None, st), "prepare")
Mp = C.mpad(M)
mean = torch.zeros(N, R, dtype=torch.float64, device=dev); var = torch.zeros_like(mean); F = torch.zeros_like(mean)
A = torch.zeros(N, Mp, dtype=torch.float64, device=dev)
U = torch.zeros(lib.dgp_usave_bytes(ctypes.byref(d), N) // 8, dtype=torch.float64, device=dev)
Ks = torch.zeros(N, Mp, dtype=torch.float64, device=dev) if a.kind == "rbf" else None
gm = t(rng.randn(N, R)); gv = t(rng.randn(N, R)); gF = t(rng.randn(N, R)); gX = torch.zeros(N, D, dtype=torch.float64, device=dev)
def fwd(): C.check(lib.dgp_cond_forward(ctypes.byref(d), C.ptr(ws), C.ptr(X), N, N, None, 1, None, 0, 0, 0, C.ptr(mean), C.ptr(var), C.ptr(F), R, C.ptr(A), C.ptr(U), C.ptr(Ks), st), "fwd")
def fwd_nosave(): C.check(lib.dgp_cond_forward(ctypes.byref(d), C.ptr(ws), C.ptr(X), N, N, None, 1, None, 0, 0, 0, C.ptr(mean), C.ptr(var), C.ptr(F), R, None, None, None, st), "fwd")
def bwd(): C.check(lib.dgp_cond_backward(ctypes.byref(d), C.ptr(ws), C.ptr(X), N, N, C.ptr(A), C.ptr(U), C.ptr(Ks), C.ptr(mean), C.ptr(var), C.ptr(F), R, C.ptr(gm), C.ptr(gv), C.ptr(gF), R, C.ptr(gX), 0, 0, st), "bwd")
def gram(): C.check(lib.dgp_cond_backward_gram(ctypes.byref(d), C.ptr(ws), C.ptr(A), N, 0, st), "gram")
def prep(): C.check(lib.dgp_cond_prepare(ctypes.byref(d), C.ptr(ws), wsb, 1, None, st), "prepare")
gZ = torch.zeros(M, D, dtype=torch.float64, device=dev); gth = torch.zeros(2 + D, dtype=torch.float64, device=dev)
gqm = torch.zeros(M, R, dtype=torch.float64, device=dev); gqs = torch.zeros(R, M, M, dtype=torch.float64, device=dev)
def tail(): C.check(lib.dgp_cond_backward_params(ctypes.byref(d), C.ptr(ws), 1.0, C.ptr(gZ), C.ptr(gth), C.ptr(gqm), C.ptr(gqs), None, None, st), "tail")
fl = {"fwd_nosave": 0, "fwd": N * (2*M*D + M*M + 2*M*R + 2*M + R*(M*M + 2*M)), "bwd": N * (R*M*M + M*M + 2*M*R + 6*M*D), "gram": N * (R + 1) * M * M, "prep": 0, "tail": 0}
out = []
fl["fwd_nosave"] = fl["fwd"]
for name, fn in (("fwd", fwd), ("fwd_nosave", fwd_nosave), ("bwd", bwd), ("gram", gram), ("prep", prep), ("tail", tail)):
    fn(); fn(); torch.cuda.synchronize(); ts = []
    for _ in range(a.reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    out.append(f"{name} {min(ts):.3f}ms ({fl[name] / min(ts) * 1e-9:.1f}TF)")
print(os.environ.get("DGP_B200_LIB", "default").split("/")[-1], f"fwd{a.force_fwd} bwd{a.force_bwd} dbg{a.debug} stg{a.stages}", f"M={M} D={D} R={R} N={N}:", " | ".join(out), flush=True)
if os.environ.get("DGP_PHASES"):
    def show(title, names):
        ph = ws[256:256 + 96].view(torch.float64).cpu().numpy()
        tot = ph[:len(names)].sum()
        print(title, {n: f"{v:.0f} ({100 * v / tot:.1f}%)" for n, v in zip(names, ph)}, flush=True)
    fwd(); torch.cuda.synchronize()
    show("fwd phases (cycles of CTA 0, all its tiles):", ["top-sync", "setup", "K", "T", "E1", "U+E2"])
    bwd(); torch.cuda.synchronize()
    show("bwd phases (cycles of CTA 0, all its tiles):", ["top-sync", "prologue", "-", "dA-loop", "dK", "store+dqmu", "store-wait", "Gr2", "T1/T2z", "finish", "dA-epilogue", "dA-entry-sync"])
