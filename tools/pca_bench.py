"""Times dgp_pca_weights (GPU PCA behind find_weights) against numpy's LAPACK SVD on the host and reports the Gram
pass against its roofline (HBM: 8 N D bytes; fp64 tensor pipe: N D (D + 8) flops on the lower 8x8 tiles)."""
import argparse, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgplib_b200 import _cabi as C

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=1000000); ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--host-svd", action="store_true")
a = ap.parse_args()
lib = C.lib()
for D in (8, 16, 32, 64):
    rng = np.random.RandomState(D)
    Xh = rng.randn(a.N, D) * np.linspace(3, 0.5, D)
    X = torch.tensor(Xh, device="cuda"); R = max(1, D // 2)
    wsb = lib.dgp_pca_ws_bytes(a.N, D); ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
    W = torch.zeros(D, R, dtype=torch.float64, device="cuda"); st = C.stream_ptr()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(a.reps + 2):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); C.check(lib.dgp_pca_weights(C.ptr(X), a.N, D, D, R, C.ptr(ws), wsb, C.ptr(W), None, st), "pca"); e1.record()
        e1.synchronize(); ts.append(e0.elapsed_time(e1))
    t = min(ts[2:])
    line = f"N={a.N} D={D}: gram+eig {t:.3f} ms; if all Gram: {8 * a.N * D / t * 1e-6:.0f} GB/s, {a.N * D * (D + 8) / t * 1e-9:.2f} TFLOP/s"
    if a.host_svd:
        t0 = time.perf_counter(); np.linalg.svd(Xh, full_matrices=False); line += f"; numpy svd {1e3 * (time.perf_counter() - t0):.0f} ms"
    print(line, flush=True)
