"""Two GPUs, one process each, NCCL: the sharded evaluation through the public API (DSDGP.distribute + elbo_and_grad /
AdamOptimizer) reproduces the unsharded evaluation of the same global minibatch on one GPU.  Skipped on boxes with fewer
than two GPUs (the CPU counterpart of the scheme is tests/test_dist_gloo.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _model():
    from _models import build_dsdgp
    return build_dsdgp(L=3, M=40, D=4, N=601, S=3, minibatch=151)       # odd sizes: uneven shards


def _worker(rank, world, port, out, graph, overlap=True):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    from dgplib_b200.parallel import nccl_options
    dist.init_process_group("nccl", rank=rank, world_size=world, pg_options=nccl_options(world))
    model, X, Y = _model()
    model.compile(f"cuda:{rank}")
    model.use_cuda_graph = graph
    model.overlap_all_reduce = overlap
    model.distribute()
    res = []
    for _ in range(3):                                                   # three consecutive minibatches
        v, g = model.elbo_and_grad()
        res.append([v.detach().cpu().clone()] + [x.detach().cpu().clone() for x in g])
    if rank == 0:
        torch.save(res, out)
    dist.barrier()
    model.release_graphs()          # graphs that captured NCCL collectives go before the communicator
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("graph,overlap", [(False, True), (True, True), (False, False), (True, False)])
def test_two_gpu_nccl_matches_single_gpu(cuda, tmp_path, graph, overlap):
    """overlap: per-layer buckets on NCCL's stream behind each layer's parameter tail (captured inside the graph when
    graph is on); otherwise one all-reduce of the whole flat buffer after the last kernel / after the replay."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = str(tmp_path / "flat.pt")
    mp.spawn(_worker, args=(2, _free_port(), out, graph, overlap), nprocs=2, join=True)
    sharded = torch.load(out)
    model, X, Y = _model()
    model.compile("cuda:0")
    for step in range(3):
        v, g = model.elbo_and_grad()
        ref = [v.detach().cpu()] + [x.detach().cpu() for x in g]
        assert len(ref) == len(sharded[step])
        for a, b in zip(sharded[step], ref):
            denom = max(1e-300, float(b.abs().max()))
            assert float((a - b).abs().max()) / denom < 1e-11


def _train_worker(rank, world, port, out, graph):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    from dgplib_b200.parallel import nccl_options
    dist.init_process_group("nccl", rank=rank, world_size=world, pg_options=nccl_options(world))
    from dgplib_b200.training import AdamOptimizer
    model, X, Y = _model()
    model.compile(f"cuda:{rank}")
    model.use_cuda_graph = graph
    model.distribute()
    opt = AdamOptimizer(0.01)
    hist = opt.minimize(model, maxiter=5)
    if rank == 0:
        torch.save(([float(h) for h in hist], [p.detach().cpu().clone() for p in model.parameters()]), out)
    dist.barrier()
    opt.release_graphs()
    model.release_graphs()
    dist.destroy_process_group()


@pytest.mark.parametrize("graph", [False, True])
def test_two_gpu_training_matches_single_gpu(cuda, tmp_path, graph):
    """Five Adam steps with the minibatch rows split over two GPUs (NCCL buckets captured inside the step graph when graph is
    on) against the same five steps on one GPU: ELBO history and every parameter."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from dgplib_b200.training import AdamOptimizer
    out = str(tmp_path / "train.pt")
    mp.spawn(_train_worker, args=(2, _free_port(), out, graph), nprocs=2, join=True)
    hist2, params2 = torch.load(out)
    model, X, Y = _model()
    model.compile("cuda:0")
    hist1 = [float(h) for h in AdamOptimizer(0.01).minimize(model, maxiter=5)]
    assert max(abs(a - b) / abs(b) for a, b in zip(hist2, hist1)) < 1e-10
    for a, b in zip(params2, [p.detach().cpu() for p in model.parameters()]):
        assert float((a - b).abs().max()) <= 1e-9 * max(1.0, float(b.abs().max()))
