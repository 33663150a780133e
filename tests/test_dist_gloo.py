"""world_size-2 gloo test (CPU) of the data-parallel scheme: each rank evaluates ELBO + gradient on its row shard with
the KL term weighted 1/W, one all-reduce(sum) of the flat buffer reproduces the unsharded evaluation.  The arithmetic
on each rank is the oracle (the CUDA engine needs a GPU); the sharding helpers are the product's own."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dgplib_b200.parallel import shard_range, kl_weight, philox_row_id


def _model():
    from oracle import dsdgp_oracle as O  # noqa: F401
    rng = np.random.RandomState(0)
    T64 = torch.float64
    t = lambda x: torch.tensor(np.array(x, dtype=np.float64), dtype=T64)
    layers = []
    for l, (D, R) in enumerate([(2, 2), (2, 1)]):
        kern = {"kind": "rbf", "input_dim": D, "variance": t(1.1), "lengthscales": t(1.3), "white": t(1e-3)}
        layers.append({"groups": [{"kernel": kern, "Z": t(rng.randn(6, D))}], "q_mu": t(0.2 * rng.randn(6, R)),
                       "q_sqrt": t(np.eye(6)[None] + 0.1 * np.tril(rng.randn(R, 6, 6))),
                       "mean": {"A": t(np.eye(2)), "b": t(np.zeros(2))} if l == 0 else None})
    lik = {"kind": "gaussian", "variance": t(0.3)}
    X, Y = t(rng.randn(10, 2)), t(rng.randn(10, 1))
    eps = [t(rng.randn(3, 10, 2)), t(rng.randn(3, 10, 1))]
    return layers, lik, X, Y, eps


def _sharded_flat(layers, lik, X, Y, eps, rank, world, num_data):
    """What one rank contributes: [elbo_local | grads_local] with KL weighted 1/W (engine.Engine.elbo's contract)."""
    from oracle import dsdgp_oracle as O
    lo, hi = shard_range(X.shape[0], rank, world)
    lv = O.leaves(layers, lik)
    for p in lv.values():
        p.requires_grad_(True)
    _, Fmeans, Fvars = O.propagate(layers, X[lo:hi], 3, [e[:, lo:hi] for e in eps])
    ve = torch.stack([O.variational_expectations(lik, Fmeans[-1][s], Fvars[-1][s], Y[lo:hi]) for s in range(3)])
    scale = float(num_data) / float(X.shape[0])           # GLOBAL minibatch size in the denominator
    val = ve.mean(0).sum() * scale - kl_weight(world) * sum(O.layer_kl(l) for l in layers)
    grads = torch.autograd.grad(val, list(lv.values()), allow_unused=True)
    for p in lv.values():
        p.requires_grad_(False)
    return torch.cat([val.detach().reshape(1)] + [torch.zeros_like(p).reshape(-1) if g is None else g.reshape(-1)
                                                   for p, g in zip(lv.values(), grads)])


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    layers, lik, X, Y, eps = _model()
    flat = _sharded_flat(layers, lik, X, Y, eps, rank, world, num_data=40)
    dist.all_reduce(flat)
    if rank == 0:
        torch.save(flat, out)
    dist.destroy_process_group()


def _bucket_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dgplib_b200.parallel import BucketReducer
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(1000, dtype=torch.float64, generator=g)
    whole = flat.clone()
    red = BucketReducer(dist)
    for lo, hi in [(700, 1000), (300, 700), (0, 300)]:        # last layer (+ likelihood) first, first layer (+ ELBO) last
        red.bucket(flat[lo:hi])
    red.finish()
    red(whole)                                                 # the single-collective path
    if rank == 0:
        torch.save((flat, whole), out)
    dist.destroy_process_group()


def test_bucketed_all_reduce_equals_the_single_collective(tmp_path):
    out = str(tmp_path / "b.pt")
    mp.spawn(_bucket_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    flat, whole = torch.load(out)
    assert torch.equal(flat, whole)
    want = sum(torch.randn(1000, dtype=torch.float64, generator=torch.Generator().manual_seed(100 + r)) for r in range(2))
    assert torch.equal(flat, want)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_all_reduce_reproduces_the_unsharded_evaluation(tmp_path):
    from oracle import dsdgp_oracle as O
    out = str(tmp_path / "flat.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    layers, lik, X, Y, eps = _model()
    val, g = O.elbo_and_grad(layers, lik, X, Y, 3, eps, num_data=40)
    want = torch.cat([val.reshape(1)] + [v.reshape(-1) for v in g.values()])
    assert float((got - want).abs().max()) < 1e-10 * float(want.abs().max())


def test_shard_helpers():
    for B, W in [(10, 2), (10000, 8), (7, 4), (80000, 8)]:
        ranges = [shard_range(B, r, W) for r in range(W)]
        assert ranges[0][0] == 0 and ranges[-1][1] == B
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    assert kl_weight(8) * 8 == 1.0
    # the Philox key of (sample s, global row b) does not depend on how rows are sharded
    B = 12
    full = {(s, b): philox_row_id(0, s, b, B) for s in range(3) for b in range(B)}
    for W in (2, 3, 4):
        for r in range(W):
            lo, hi = shard_range(B, r, W)
            for s in range(3):
                for b in range(lo, hi):
                    assert philox_row_id(lo, s, b - lo, B) == full[(s, b)]


def test_nccl_options_cta_budget_by_world_size():
    """parallel.nccl_options: four CTAs per collective for two-rank groups (measured), NCCL's own choice otherwise, an
    explicit request always honoured."""
    import torch.distributed as dist
    from dgplib_b200.parallel import nccl_options
    if not hasattr(dist, "ProcessGroupNCCL"):
        pytest.skip("torch built without NCCL")
    default = dist.ProcessGroupNCCL.Options().config.max_ctas
    assert nccl_options(2).config.max_ctas == 4 and nccl_options(2).config.min_ctas == 1
    assert nccl_options(4).config.max_ctas == default and nccl_options(8).config.max_ctas == default
    assert nccl_options(8, max_ctas=16).config.max_ctas == 16
    assert nccl_options().config.max_ctas == default
