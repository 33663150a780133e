"""cuBLAS DGEMM throughput on this B200 (torch.matmul fp64): burst (best of 10) and sustained (4 s loop).
Roofline-denominator measurement only; not on the product path."""
import json, time, torch
def run(n):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(3): a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); e1.synchronize(); best = min(best, e0.elapsed_time(e1))
    burst = 2 * n**3 / best * 1e-9
    t0 = time.time(); cnt = 0
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
    while time.time() - t0 < 4.0:
        for _ in range(5): a @ b
        cnt += 5; torch.cuda.synchronize()
    e1.record(); e1.synchronize()
    sus = 2 * n**3 * cnt / e0.elapsed_time(e1) * 1e-9
    return burst, sus
out = {}
for n in (4096, 8192):
    b, s = run(n); out[f"dgemm_{n}"] = {"burst_tflops": b, "sustained_tflops": s}; print(n, b, s, flush=True)
json.dump(out, open("gpurun_out/fp64_peak_torch.json", "w"), indent=1)
