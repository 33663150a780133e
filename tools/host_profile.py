"""Host-side profile of the public per-step API on a small, latency-bound model (BASELINE configs[0]): where the Python
time of DSDGP.elbo_and_grad() goes.  Usage: python tools/host_profile.py [workload] [steps]"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "C1"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
w = bench.WORKLOADS[name]
model = bench.build_model(w, w["B"])
model.compile("cuda:0")
model.use_cuda_graph = True
for _ in range(20):
    model.elbo_and_grad()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    model.elbo_and_grad()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"{name}: host enqueue {1e3 * (t1 - t0) / steps:.4f} ms/step, with final sync {1e3 * (t2 - t0) / steps:.4f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(steps):
    model.elbo_and_grad()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr).sort_stats("tottime")
st.print_stats(22)
st.print_callers("parameters")
st.print_callers("raw_parameters")
