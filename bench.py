#!/usr/bin/env python
"""Benchmark of the DSDGP ELBO+gradient hot path (BASELINE.json metric): rows/s = B*S*evals/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload T|C2|C1|C3s|C4|C5] [--impl ours|reference]

N > 1 is launched by the driver with torch.distributed.run (one rank per GPU, NCCL).  Prints ONE JSON line (rank 0).
`value`   : device-timed throughput, inputs resident in HBM.
`e2e`     : the same metric through the public API (DSDGP.elbo_and_grad) with the minibatch copied from pinned host
            memory and the ELBO read back every step, Cholesky status check left at its default (every step).
`roofline`: dominant kernel's algorithmic flops / CUDA-event time vs the measured fp64 DMMA peak, plus three step-level
            fractions (executed flops, SURVEY 8(d) algorithmic flops with the first layer evaluated once per data row,
            and the reference's own flop count with S copies of the first layer).
`cpu_baseline` / --impl reference: the oracle (op-for-op restatement of the reference graph in PyTorch-CPU fp64; TF 1.8 /
            GPflow 1.2 cannot be installed here) on the box's host cores, bounded sample.
`workloads`: condensed records of the other BASELINE configurations (default single-GPU invocation of T only).
`strong`  : fixed-size global minibatches split over the N ranks (configs[2]'s 80 k rows and T's 10 k rows).
`shard_parity_rel_err` (N > 1): the all-reduced sharded ELBO + gradient against rank 0's unsharded evaluation of the same
            global minibatch with the same Philox seed.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: L, M, D, N (dataset rows), B (minibatch rows per GPU), S, kind
    "T":   dict(L=3, M=256, D=8, N=100_000, B=10_000, S=10, kind="rbf",
                desc="3-layer DSDGP RBF M=256 D=8 N=100k minibatch 10k S=10 (north_star target line = configs[1] with M=256)"),
    "C2":  dict(L=3, M=128, D=8, N=100_000, B=10_000, S=10, kind="rbf",
                desc="configs[1]: 3-layer DSDGP RBF M=128 D=8 N=100k minibatch 10k S=10"),
    "C1":  dict(L=2, M=50, D=1, N=1000, B=1000, S=1, kind="rbf", desc="configs[0]: 2-layer RBF M=50 D=1 N=1000 S=1"),
    "C3s": dict(L=5, M=500, D=16, N=1_000_000, B=10_000, S=16, kind="matern52",
                desc="configs[2] per-GPU shard: 5-layer Matern52 M=500 D=16 minibatch 80k/8 S=16"),
    "C3":  dict(L=5, M=500, D=16, N=1_000_000, B=80_000, S=16, kind="matern52", chunk=10_000,
                desc="configs[2]: 5-layer Matern52 M=500 D=16 N=1M minibatch 80k S=16 (global minibatch split over the GPUs)"),
    "C4":  dict(L=2, M=256, D=8, N=200_000, B=10_000, S=10, kind="rbf", tasks=8,
                desc="configs[3]: multitask DSDGP, MultikernelInputLayer (8 RBF kernels) + SwitchedKernel output layer, "
                     "SwitchedLikelihood, M=256, 8 tasks, N=200k, minibatch 10k S=10"),
    "C5":  dict(L=3, M=1024, D=8, N=100_000, B=10_000, S=100, kind="rbf", predict=True,
                desc="configs[4] per-GPU slice: predict_y sweep, 3-layer M=1024, S=100, 10k test points per step "
                     "(forward only)"),
}
FP64_PEAK_TFLOPS = 37.1   # profiles/fp64_peak_r01.txt: DMMA.8x8x4 issue-rate microbenchmark on this pool's B200
                          # (cuBLAS DGEMM 8192^3: 35.5).  MEASURED_PEAKS.json holds no fp64 figure.
PEAK_SOURCE = ("fp64 DMMA.8x8x4 issue-rate microbenchmark on this pool's B200 (profiles/fp64_peak_r02.txt); "
               "MEASURED_PEAKS.json has no fp64 entry (its bf16 tcgen05 peak does not apply: there is no fp64 tcgen05 kind)")


def flops_per_row_layer(M, D_in, R):
    """SURVEY.md 8(d) forward flop convention per row-layer; ELBO+grad = 3x."""
    return 2 * M * D_in + M * M + 2 * M * R + 2 * M + R * (M * M + 2 * M)


def build_model(w, global_rows):
    import dgplib_b200 as dg
    from dgplib_b200.cascade import Sequential
    from dgplib_b200.kernels import RBF, Matern52, White
    from dgplib_b200.layers import InputLayer, HiddenLayer, OutputLayer
    from dgplib_b200.likelihoods import Gaussian
    from dgplib_b200.mean_functions import Linear
    L, M, D, N, S = w["L"], w["M"], w["D"], w["N"], w["S"]
    if w.get("tasks"):
        return build_multitask_model(w, global_rows)
    rng = np.random.RandomState(0)
    X = rng.randn(N, D)
    Y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.randn(N, 1)
    if D == 1:
        X = rng.uniform(0, 1, (N, 1))
    Z = X[rng.permutation(N)[:M]].copy()
    dims = [D] * L + [1]
    cls_k = RBF if w["kind"] == "rbf" else Matern52
    layers = []
    for l in range(L):
        cls = InputLayer if l == 0 else (OutputLayer if l == L - 1 else HiddenLayer)
        kern = cls_k(dims[l], variance=1.0, lengthscales=float(np.sqrt(dims[l]))) + White(dims[l], variance=1e-5)
        mf = Linear(A=np.zeros((dims[l], dims[l + 1]))) if l < L - 1 else None
        layers.append(cls(dims[l], dims[l + 1], M, kern, mean_function=mf))
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        model = dg.DSDGP(X, Y, Z, Sequential(layers), Gaussian(variance=0.1), minibatch_size=min(global_rows, N),
                         num_samples=S, seed=1234)
    rs = np.random.RandomState(1)
    for layer in model.layers.layers:
        Mq, R = layer.q_mu.shape
        layer.q_mu.assign(0.1 * rs.randn(Mq, R))
        layer.q_sqrt.assign(np.eye(Mq)[None] + 0.1 * np.tril(rs.randn(R, Mq, Mq)))
    return model


def build_multitask_model(w, global_rows):
    """BASELINE configs[3]: task id column appended to X, Z, Y; MultikernelInputLayer with T RBF kernels
    (lengthscale sqrt(D) (0.8 + 0.1 k)), SwitchedKernel output layer, SwitchedLikelihood of T Gaussians."""
    import contextlib, io
    import dgplib_b200 as dg
    from dgplib_b200.cascade import MultitaskSequential
    from dgplib_b200.kernels import RBF, White
    from dgplib_b200.layers import OutputLayer
    from dgplib_b200.likelihoods import Gaussian, SwitchedLikelihood
    from dgplib_b200.mean_functions import Linear
    from dgplib_b200.multikernel_layers import MultikernelInputLayer
    from dgplib_b200.specialized_kernels import SwitchedKernel
    M, D, N, S, T = w["M"], w["D"], w["N"], w["S"], w["tasks"]
    rng = np.random.RandomState(0)
    Xv = rng.randn(N, D)
    tid = rng.randint(0, T, (N, 1)).astype(np.float64)
    X = np.hstack([Xv, tid])
    Y = np.hstack([np.sin(Xv.sum(1, keepdims=True)) + 0.1 * rng.randn(N, 1), tid])
    Z = X[rng.permutation(N)[:M]].copy()
    il = MultikernelInputLayer(D, T, M, [RBF(D, lengthscales=float(np.sqrt(D) * (0.8 + 0.1 * k))) + White(D, 1e-5)
                                         for k in range(T)], share_Z=False,
                               mean_function=Linear(A=np.zeros((D + 1, T))), multitask=True)
    ol = OutputLayer(T, 1, M, SwitchedKernel([RBF(T, lengthscales=float(np.sqrt(T))) + White(T, 1e-5) for _ in range(T)], T),
                     multitask=True)
    with contextlib.redirect_stdout(io.StringIO()):
        model = dg.MultitaskDSDGP(X, Y, Z, MultitaskSequential([il, ol]),
                                  SwitchedLikelihood([Gaussian(0.1) for _ in range(T)]), num_latent=1,
                                  minibatch_size=min(global_rows, N), num_samples=S, seed=1234)
    rs = np.random.RandomState(1)
    for layer in model.layers.layers:
        Mq, R = layer.q_mu.shape
        layer.q_mu.assign(0.1 * rs.randn(Mq, R))
        layer.q_sqrt.assign(np.eye(Mq)[None] + 0.1 * np.tril(rs.randn(R, Mq, Mq)))
    return model


def workload_flops_per_row(w):
    if w.get("tasks"):      # MultikernelLayer: F = K (2 M D + M^2) + 2 M R + 2 M K + R (M^2 + 2 M); SURVEY.md 8(d)
        M, D, T = w["M"], w["D"], w["tasks"]
        first = T * (2 * M * D + M * M) + 2 * M * T + 2 * M * T + T * (M * M + 2 * M)
        return first + flops_per_row_layer(M, T, 1)
    dims = [w["D"]] * w["L"] + [1]
    return sum(flops_per_row_layer(w["M"], dims[l], dims[l + 1]) for l in range(w["L"]))


def make_config(name, w, world, B):
    """The workload description shared verbatim by our arm and the reference arm (the driver compares them)."""
    return {"workload": f"{name}: {w['desc']}", "layers": w["L"], "M": w["M"], "D": w["D"], "S": w["S"],
            "minibatch_rows_per_gpu": B, "global_minibatch_rows": B * max(world, 1), "kernel": w["kind"],
            "sharding": f"dp{max(world, 1)} over minibatch rows",
            "l2": "per-step working set (A and U tiles, dKuf) is >1 GB, far above the 126 MB L2; no flush needed"}


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = sorted(float(r[1]) for r in rows)
        out["sm_mhz"] = sm[len(sm) // 2]
        out["sm_max_mhz"] = float(rows[0][2])
        out["power_w_max"] = max(float(r[3]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i, nm in enumerate(names):
            if any("Active" == r[5 + i].strip() for r in rows):
                out["reasons"].append(nm)
        out["samples"] = len(rows)
        return out


def cpu_oracle_throughput(model, w, sample_rows, reps, warmup=1, threads=None):
    """Oracle on the host cores over `sample_rows` rows (x S samples): ELBO+grad (or, for the predict workload, the
    forward sweep).  Returns (rows/s from the MEAN of `reps` evaluations after `warmup`, seconds per evaluation, threads)."""
    from oracle import dsdgp_oracle as O
    from oracle.bridge import oracle_layers, oracle_lik
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    layers, lik = oracle_layers(model.layers.layers), oracle_lik(model.likelihood)
    S = w["S"]
    X = model.X[:sample_rows].clone()
    Y = model.Y[:sample_rows].clone()
    eps = [torch.randn(S, sample_rows, l["q_mu"].shape[1], dtype=torch.float64) for l in layers]
    mt = bool(w.get("tasks"))
    times = []
    for i in range(warmup + reps):
        t0 = time.perf_counter()
        if w.get("predict"):
            with torch.no_grad():
                O.predict_y(layers, lik, X, eps, S=S, multitask=mt)
        else:
            O.elbo_and_grad(layers, lik, X, Y, S, eps, num_data=w["N"], multitask=mt)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    dt = float(np.mean(times))
    return sample_rows * S / dt, dt, torch.get_num_threads()


def cpu_sample_rows(w, budget_flops, cap_rows=None):
    """Bounded sample of the workload for the CPU arm: at most `budget_flops` algorithmic flops per evaluation, never more
    than the minibatch, and small enough that the oracle's materialised [R, M, S*rows] tensors (8 R M S bytes per row and
    layer, kept for autograd) fit in host RAM."""
    per_row = (1 if w.get("predict") else 3) * workload_flops_per_row(w) * w["S"]
    rows = int(max(64, min(cap_rows or w["B"], budget_flops / per_row)))
    lta_per_row = 8 * w["L"] * max(w["D"], w.get("tasks") or 1) * w["M"] * w["S"] * 3      # LTA + A + autograd copies
    return int(max(64, min(rows, 16e9 / lta_per_row)))


def cpu_record(model, w, rows, reps, warmup, B):
    v, dt, thr = cpu_oracle_throughput(model, w, rows, reps, warmup=warmup)
    what = "test rows" if w.get("predict") else "minibatch rows"
    return {"value": v, "unit": "rows/s", "cores": thr, "kind": "port",
            "sample": f"{rows} of {B} {what} x S={w['S']} per evaluation, mean of {reps} after {warmup} warm-up "
                      f"({dt:.2f} s/eval): oracle = PyTorch-CPU fp64 restatement of the reference graph "
                      "(TF 1.8 / GPflow 1.2 not installable)"}


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from committed `ncu --set full` captures, keyed by kernel
    and shape (profiles/r02_traffic.json, written by tools/ncu_summary.py); a shape without a capture reports null."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    except (OSError, ValueError):
        return {}


def kernel_flop_table(eng, B, S):
    """Per kernel class and layer: algorithmic flops one launch set executes (sums a layer's conditionals)."""
    per_kernel, shapes = {}, {}
    for l, conds in enumerate(eng.conds):
        n_rows = B * S if l > 0 else B      # the first layer is evaluated once per data row (sample sharing)
        f = b = g = 0
        for c in conds:
            M, Din, R = c.desc.M, c.desc.D, c.desc.R
            f += n_rows * flops_per_row_layer(M, Din, R)
            b += n_rows * (R * M * M + M * M + 2 * M * R + 6 * M * Din)   # dA (triangular Lq_r U_r) + Lm^-T + d q_mu + kernel bwd
            g += n_rows * (R + 1) * M * M                                  # G_r (sym) + P (tril)
        per_kernel[f"fwd.l{l}"], per_kernel[f"bwd.l{l}"], per_kernel[f"gram.l{l}"] = f, b, g
        c0 = conds[0].desc
        for k, nm in (("fwd", "k_cond_fwd"), ("bwd", "k_cond_bwd"), ("gram", "k_gram")):
            shapes[f"{k}.l{l}"] = f"{nm}|M{c0.M},D{c0.D},R{c0.R},N{n_rows}"
    return per_kernel, shapes


def step_flop_counts(w, eng, B, S):
    """(executed, SURVEY-8(d) algorithmic with the first layer evaluated once per data row, reference count with S copies
    of the first layer) flops of one ELBO+grad evaluation of B data rows."""
    per_kernel, _ = kernel_flop_table(eng, B, S)
    executed = sum(per_kernel.values())
    alg = 0
    for l, conds in enumerate(eng.conds):
        n_rows = B * S if l > 0 else B
        alg += 3 * n_rows * sum(flops_per_row_layer(c.desc.M, c.desc.D, c.desc.R) for c in conds)
    reference = 3 * workload_flops_per_row(w) * B * S
    return executed, alg, reference


def run_predict(args, name, w, model, eng, rank, world, local_rank, dist, sync_all, config, with_cpu):
    """Forward-only sweep (BASELINE configs[4]): rows/s = N* S / t; every rank predicts its own slice of test points."""
    from dgplib_b200 import _cabi
    S, B = w["S"], w["B"]
    dev = eng.device
    metric = "DSDGP predict_y rows/s (N* x S / s, forward only)"
    rng = np.random.RandomState(4 + rank)
    Xs_host = torch.from_numpy(rng.randn(B, w["D"])).pin_memory()
    Xs = Xs_host.to(dev)
    chunk = 2048

    def step_dev(i):
        # the model's own parameters: the parameter stage (Kuu, Cholesky, inverse at M = 1024) is prepared once per sweep
        return eng.propagate(Xs, S, seed=100 + i, chunk_rows=chunk, last_only=True)

    for i in range(max(3, args.warmup)):
        step_dev(i)
    sync_all()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(1.5)
    gc.collect(); gc.disable()
    eng.timers = {}
    l0 = _cabi.launch_count()
    sync_all()               # all ranks enter the timed region together (rank 0 alone started the clock sampler)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_dev(1000 + i)
    e1.record()
    sync_all()
    gc.enable()
    launches = _cabi.launch_count() - l0
    ms = e0.elapsed_time(e1)
    timers = eng.timer_summary()
    eng.timers = None
    clocks = sampler.stop() if sampler is not None else None
    # end to end: pinned host test points -> device, predict_y, mean/var back to the host every step
    model.predict_chunk_rows = chunk
    model.predict_y(Xs_host, num_samples=S)
    sync_all()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        mu, var = model.predict_y(Xs_host, num_samples=S, seed=2000 + i)
    t1.record()
    sync_all()
    ms_e2e = t0.elapsed_time(t1)
    if dist is not None:
        tm = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(tm[0]), float(tm[1])
    rows = B * S * world
    per_kernel = {}
    for l, conds in enumerate(eng.conds):
        n_rows = B * S if l > 0 else B
        per_kernel[f"fwd.l{l}"] = sum(n_rows * flops_per_row_layer(c.desc.M, c.desc.D, c.desc.R) for c in conds)
    dom = max((k for k in timers if k in per_kernel), key=lambda k: timers[k][1])
    n_l, ms_l = timers[dom]
    achieved = per_kernel[dom] * args.steps / (ms_l * 1e-3) / 1e12
    step_alg = sum(per_kernel.values())
    line = {"metric": metric, "value": rows * args.steps / (ms * 1e-3), "unit": "rows/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "e2e": {"value": rows * args.steps / (ms_e2e * 1e-3), "unit": "rows/s", "h2d_bytes_per_step": B * w["D"] * 8,
                    "d2h_bytes_per_step": 2 * S * B * 8, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": FP64_PEAK_TFLOPS,
                         "unit": "TFLOP/s", "frac": achieved / FP64_PEAK_TFLOPS, "traffic": None,
                         "peak_source": PEAK_SOURCE,
                         "step_frac_executed": step_alg * args.steps / (ms * 1e-3) / 1e12 / FP64_PEAK_TFLOPS,
                         "kernel_ms": {k: round(v[1] / max(args.steps, 1), 4) for k, v in sorted(timers.items())}}}
    if with_cpu:
        rows_cpu = cpu_sample_rows(w, args.cpu_flops)
        line["cpu_baseline"] = cpu_record(model, w, rows_cpu, 2, 1, B)
    return line


def run_elbo(args, name, w, rank, world, local_rank, dist, sync_all, with_cpu, with_train=True, with_parity=False):
    """One ELBO+grad workload on this rank's shard: device-resident leg, per-kernel timing pass, end-to-end leg, training
    step; returns the JSON line as a dict (meaningful on rank 0)."""
    from dgplib_b200 import _cabi
    S, B = w["S"], w["B"]
    strong = bool(w.get("strong"))
    Bg = B if strong else B * world                 # strong: the global minibatch is fixed and split over the ranks
    config = make_config(name, w, world, B // world if strong else B)
    if strong:
        config["global_minibatch_rows"] = B
    model = build_model(w, Bg)
    model.compile(f"cuda:{local_rank}")
    model.chunk_rows = w.get("chunk")
    if dist is not None:
        model.distribute()
    eng = model.engine
    dev = eng.device
    if w.get("predict"):
        line = run_predict(args, name, w, model, eng, rank, world, local_rank, dist, sync_all, config, with_cpu)
        del model, eng
        gc.collect(); torch.cuda.empty_cache()
        return line
    metric = "DSDGP ELBO+grad rows/s (B*S*evals/s)"
    from dgplib_b200.parallel import shard_range
    lo, hi = shard_range(Bg, rank, world)
    Bl = hi - lo
    Xb, Yb = model.X[:Bg], model.Y[:Bg]
    Xd, Yd = Xb[lo:hi].to(dev), Yb[lo:hi].to(dev)
    from dgplib_b200.parallel import BucketReducer
    # per-layer gradient buckets, all-reduced on NCCL's stream as soon as each layer's parameter tail has finished and
    # captured inside the step's CUDA graph
    ar = BucketReducer(dist) if dist is not None else None
    chunk = w.get("chunk")

    graphed = None
    if not args.no_graph:
        from dgplib_b200.engine import GraphedElbo
        # live parameters: the graph reads the model's own storage, the hyper-parameter transforms are part of the capture
        graphed = GraphedElbo(eng, None, Xd, Yd, S, num_data=w["N"], row_offset=lo, global_batch=Bg, world_size=world,
                              base_seed=1234, chunk_rows=chunk, all_reduce=ar)
        model.use_cuda_graph = True

    def step_eager(i):
        return eng.elbo(eng.native_params(), Xd, Yd, S, eps=None, seed=1234 + i, num_data=w["N"], need_grad=True,
                        row_offset=lo, global_batch=Bg, world_size=world, all_reduce=ar, check=False, chunk_rows=chunk)

    def step_dev(i):
        if graphed is not None:
            return graphed.run(None, Xd, Yd, 1234 + i)
        return step_eager(i)

    l0 = _cabi.launch_count()
    step_eager(0)
    launches_per_step = _cabi.launch_count() - l0
    for i in range(max(3, args.warmup)):
        step_dev(i)
    eng.check_status()
    sync_all()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(1.5)          # let nvidia-smi finish its NVML start-up before the timed region
    gc.collect()
    gc.disable()
    sync_all()               # barrier + synchronize right in front of the timed region: rank 0 alone started the clock sampler,
                             # and a rank that enters late makes the others wait inside their first collective
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step_evs = []
    host_t0 = time.perf_counter()
    for i in range(args.steps):
        val, _ = step_dev(100 + i)
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        step_evs.append(ev)
    e1.record()
    host_ms = (time.perf_counter() - host_t0) * 1e3 / args.steps     # host time to enqueue one step (no sync inside)
    sync_all()
    ms = e0.elapsed_time(e1)
    ms_steps = [round(a.elapsed_time(b), 3) for a, b in zip([e0] + step_evs[:-1], step_evs)]
    elbo_val = float(val)
    # per-kernel CUDA-event times: the same K steps once more, launched eagerly with events around every C-ABI call, with
    # the parameter stages, Gram reductions and parameter tails all on the main stream: beside another kernel an event pair
    # would time the wait for SMs, not the kernel
    eng.timers = {}
    side_gram, side_par = eng.gram_on_side_stream, eng.params_on_side_streams
    eng.gram_on_side_stream = eng.params_on_side_streams = False
    for i in range(args.steps):
        step_eager(100 + i)
    sync_all()
    eng.gram_on_side_stream, eng.params_on_side_streams = side_gram, side_par
    gc.enable()
    timers = eng.timer_summary()
    eng.timers = None
    clocks = sampler.stop() if sampler is not None else None
    launches = launches_per_step * args.steps     # kernels of the timed K steps (replayed from the captured graph)
    if dist is not None:
        tm = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm)
    rows_per_step = Bg * S
    value = rows_per_step * args.steps / (ms * 1e-3)

    # ---- sharded against unsharded: the all-reduced [elbo | gradients] buffer of one more step against rank 0's evaluation
    #      of the whole global minibatch with the same Philox seed
    parity = None
    if with_parity and dist is not None:
        step_eager(7777)
        sharded = eng._last_flat.clone()
        if rank == 0:
            Xg, Yg = Xb.to(dev), Yb.to(dev)
            eng.elbo(eng.native_params(), Xg, Yg, S, eps=None, seed=1234 + 7777, num_data=w["N"], need_grad=True, row_offset=0,
                     global_batch=Bg, world_size=1, all_reduce=None, check=False, chunk_rows=chunk or 10_000)
            whole = eng._last_flat
            parity = float((sharded - whole).abs().max() / whole.abs().max())
            del Xg, Yg
        sync_all()

    # ---- end-to-end leg: public API, pinned host minibatch -> device every step, ELBO read back every step
    # (the read is an async copy into pinned memory, consumed one step later; the Cholesky status check stays at its
    # default, every step: the status words travel the same way and are looked at when the next step is queued)
    host_elbo = torch.empty(args.steps + 2, dtype=torch.float64).pin_memory()
    for i in range(2):
        v, _ = model.elbo_and_grad()
        float(v)
    sync_all()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    evs = []
    for i in range(args.steps):
        v, g = model.elbo_and_grad()
        host_elbo[i:i + 1].copy_(v.reshape(1), non_blocking=True)      # D2H read of the step's result
        evs.append(torch.cuda.current_stream().record_event())
        if i > 0:
            evs[i - 1].synchronize()
            assert np.isfinite(float(host_elbo[i - 1]))
    evs[-1].synchronize()
    assert np.isfinite(float(host_elbo[args.steps - 1]))
    model.engine.poll_status()          # the last step's Cholesky status (every earlier one was looked at one step later)
    t1.record()
    sync_all()
    ms_e2e = t0.elapsed_time(t1)
    if dist is not None:
        tm = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms_e2e = float(tm)
    e2e_value = rows_per_step * args.steps / (ms_e2e * 1e-3)
    # ---- training step around the path: TF-Adam on the unconstrained parameters, same minibatches
    train = None
    if with_train:
        from dgplib_b200.training import AdamOptimizer
        opt = AdamOptimizer(1e-3)
        opt.minimize(model, maxiter=3)
        sync_all()
        l0 = _cabi.launch_count()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        opt.minimize(model, maxiter=args.steps)
        t1.record()
        sync_all()
        ms_train = t0.elapsed_time(t1) / args.steps
        train = {"ms_per_step": ms_train, "rows_per_s": rows_per_step / (ms_train * 1e-3),
                 "what": "AdamOptimizer.minimize: dgp_transform_params + fused ELBO/grad + dgp_adam_step (softplus chain + "
                         "tf.train Adam on the flat unconstrained buffer)"
                         + (" replayed as one CUDA graph per step" if model.use_cuda_graph and dist is None else "")
                         + ", minibatch from pinned host memory, Cholesky status read every step"}
    Dx, ldy = model.X.shape[1], model.Y.shape[1]
    h2d = Bl * (Dx + ldy) * 8       # per rank, the rows it copies
    d2h = 8 + 4 * sum(len(c) for c in eng.conds)      # the ELBO + one status word per conditional

    # ---- roofline of the dominant kernel (CUDA events on the launching stream, inside the timing pass above)
    per_kernel, shapes = kernel_flop_table(eng, Bl, S)
    dom = max((k for k in timers if k in per_kernel), key=lambda k: timers[k][1])
    n_l, ms_l = timers[dom]
    achieved = per_kernel[dom] * args.steps / (ms_l * 1e-3) / 1e12     # per_kernel sums a layer's conditionals
    executed, alg, reference = step_flop_counts(w, eng, Bl, S)
    tfl = lambda f: f * args.steps / (ms * 1e-3) / 1e12
    traffic = load_traffic().get(shapes[dom])
    roofline = {"bound": "tensor", "kernel": dom, "kernel_shape": shapes[dom], "achieved": achieved,
                "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP64_PEAK_TFLOPS,
                "traffic": traffic, "traffic_source": "profiles/r02_traffic.json (ncu --set full, dram__bytes_read.sum + "
                                                      "dram__bytes_write.sum per launch)" if traffic else None,
                "peak_source": PEAK_SOURCE,
                "kernel_ms": {k: round(v[1] / max(args.steps, 1), 4) for k, v in sorted(timers.items())},
                "kernel_frac": {k: round(per_kernel[k] * args.steps / (timers[k][1] * 1e-3) / 1e12 / FP64_PEAK_TFLOPS, 3)
                                for k in sorted(per_kernel) if k in timers and timers[k][1] > 0},
                "kernel_timing": "CUDA events around every C-ABI call on its launching stream, K eagerly launched steps "
                                 "run right after the timed K steps (which replay one CUDA graph per step), with parameter "
                                 "stages, Gram reductions and parameter tails kept on the main stream so that no timed "
                                 "kernel overlaps another",
                # three step-level fractions of the fp64 peak (per rank), named for what they count
                "step_frac_executed": tfl(executed) / FP64_PEAK_TFLOPS,
                "step_frac_algorithmic_8d": tfl(alg) / FP64_PEAK_TFLOPS,
                "step_frac_reference_count": tfl(reference) / FP64_PEAK_TFLOPS,
                "step_flops": {"executed": executed, "algorithmic_8d_layer0_once": alg, "reference_count": reference,
                               "note": "executed = what the row kernels do (fwd F; bwd R M^2 + M^2 + ...; Gram (R+1) M^2); "
                                       "algorithmic_8d = 3 F per row-layer with the first layer on B rows; reference_count "
                                       "= 3 F with the first layer on S B rows, as the reference evaluates it"}}

    line = {"metric": metric, "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "collective": None if dist is None else "NCCL all-reduce(sum) of the flat fp64 [elbo | gradients] buffer in one "
                          "bucket per layer, issued behind that layer's parameter tail and captured inside the step graph",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "elbo": elbo_val,
            "launch": "eager" if graphed is None else "one CUDA graph per evaluation",
            "e2e": {"value": e2e_value, "unit": "rows/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "gpu_launches_per_step": int(launches_per_step), "clocks": clocks,
            "roofline": roofline, "ms_steps": ms_steps, "train": train,
            "host_enqueue_ms_per_step": round(host_ms, 3)}
    if parity is not None:
        line["shard_parity_rel_err"] = parity
    if with_cpu:
        rows = cpu_sample_rows(w, args.cpu_flops)
        line["cpu_baseline"] = cpu_record(model, w, rows, 2, 1, B)
    model.release_graphs()
    del model, eng, graphed
    gc.collect(); torch.cuda.empty_cache()
    return line


def condensed(j):
    out = {"workload": j["config"]["workload"], "metric": j["metric"], "value": j["value"], "unit": "rows/s",
           "ms_per_step": j["ms_per_step"], "scaling": j["scaling"], "e2e": j["e2e"],
           "roofline": {k: j["roofline"].get(k) for k in ("kernel", "achieved", "peak", "frac", "traffic", "kernel_ms",
                                                          "step_frac_executed", "step_frac_algorithmic_8d",
                                                          "step_frac_reference_count")},
           "clocks": j["clocks"], "gpu_launches_per_step": j.get("gpu_launches_per_step")}
    for k in ("cpu_baseline", "train", "launch", "shard_parity_rel_err"):
        if j.get(k) is not None:
            out[k] = j[k]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="T", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-flops", type=float, default=4.1e11,
                    help="algorithmic flops per CPU-baseline evaluation (bounds the sample; 4.1e11 = the whole T minibatch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the condensed records of the other BASELINE workloads / the strong-scaling legs")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if os.environ.get("DGP_BENCH_WATCHDOG"):          # debugging aid: dump every thread's stack and exit after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["DGP_BENCH_WATCHDOG"]), exit=True)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        S, B = w["S"], w["B"]
        config = make_config(args.workload, w, args.gpus, B)
        model = build_model(w, B)
        rows = cpu_sample_rows(w, args.cpu_flops)
        rec = cpu_record(model, w, rows, max(1, args.steps), max(0, args.warmup), B)
        v = rec["value"]
        metric = "DSDGP predict_y rows/s (N* x S / s, forward only)" if w.get("predict") else \
            "DSDGP ELBO+grad rows/s (B*S*evals/s)"
        line = {"impl": "reference", "metric": metric, "value": v, "unit": "rows/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * rows * S / v, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": rec,
                "e2e": {"value": v, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------------------------------------ our arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        from dgplib_b200.parallel import nccl_options
        # NCCL CTA budget by world size: the gradient buckets overlap the persistent backward kernels, which hold every SM
        # (measurements in parallel.nccl_options)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"), pg_options=nccl_options(world))

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    cpu_ok = rank == 0 and world == 1 and not args.no_cpu_baseline
    line = run_elbo(args, args.workload, w, rank, world, local_rank, dist, sync_all, with_cpu=cpu_ok,
                    with_parity=world > 1)

    if args.workload == "T" and not args.no_secondary:
        short = argparse.Namespace(**vars(args))
        short.steps, short.warmup = max(3, min(args.steps, 5)), 3
        short.cpu_flops = min(args.cpu_flops, 1.0e11)      # ~1 s per CPU evaluation for the condensed records
        if world == 1:
            # the other BASELINE configurations, condensed (each the full measurement above at fewer steps)
            recs = {}
            for nm in ("C1", "C2", "C3s", "C4", "C5"):
                try:
                    recs[nm] = condensed(run_elbo(short, nm, WORKLOADS[nm], rank, world, local_rank, dist, sync_all,
                                                  with_cpu=cpu_ok, with_train=(nm in ("C1", "C2"))))
                except Exception as exc:   # the headline line must not depend on the extra runs
                    recs[nm] = {"error": repr(exc)[:300]}
            line["workloads"] = recs
        # strong scaling: fixed global minibatches split over the ranks (the N = 1 lines are the anchors of the curves)
        strong = {}
        for nm, base, steps in (("T_10k", "T", short.steps), ("C3_80k", "C3", 2)):
            ws = dict(WORKLOADS[base], strong=True)
            sa = argparse.Namespace(**vars(short))
            sa.steps = steps
            try:
                j = run_elbo(sa, base, ws, rank, world, local_rank, dist, sync_all, with_cpu=False, with_train=False)
                strong[nm] = {"global_minibatch_rows": ws["B"], "rows_per_gpu": ws["B"] // world, "value": j["value"],
                              "unit": "rows/s", "ms_per_step": j["ms_per_step"], "e2e": j["e2e"]["value"],
                              "roofline_kernel": j["roofline"]["kernel"], "roofline_frac": j["roofline"]["frac"],
                              "row_chunks_per_step": (-(-(ws["B"] // world) // ws["chunk"]) if ws.get("chunk") else 1)}
            except Exception as exc:
                strong[nm] = {"error": repr(exc)[:300]}
        line["strong"] = strong
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        gc.collect()                   # every captured graph (they hold NCCL kernels) is gone before the communicator
        torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
