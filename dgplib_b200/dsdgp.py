"""DSDGP -- doubly stochastic deep GP model with the surface of dgplib/dsdgp.py (Salimbeni & Deisenroth 2017).

The graph the reference builds with TensorFlow (`_propagate`, `_build_predict`, `_build_likelihood`,
dgplib/dsdgp.py:84-130) is executed here by the fused sm_100a kernels behind `Engine`; gradients of the ELBO w.r.t.
every trainable parameter come from hand-written backward kernels and are exposed to torch.autograd / torch.optim as
one autograd node.  Noise draws: the reference uses unseeded tf.random_normal (dgplib/utils.py:26); here they are
Philox4x32-10 draws keyed by (seed, layer, global row, output) unless explicit `eps` tensors are injected.
"""
import numpy as np
import torch

from . import settings
from .engine import Engine, elbo_autograd, _dev_tensor
from .mean_functions import Zero
from .parallel import BucketReducer, shard_range
from .params import Parameterized


class _Minibatch:
    """gpflow.params.Minibatch(v, batch_size, seed): repeat -> shuffle(buffer=N, seed) -> batch, one batch per
    evaluation.  X and Y use the same seed so the shuffles stay aligned (dgplib/dsdgp.py:78-79)."""

    def __init__(self, n, batch_size, seed):
        self.n, self.batch_size = n, int(batch_size)
        self.rng = np.random.RandomState(seed)
        self.perm = self.rng.permutation(n)
        self.pos = 0

    def next_indices(self):
        idx = []
        need = self.batch_size
        while need > 0:
            if self.pos == self.n:
                self.perm = self.rng.permutation(self.n)
                self.pos = 0
            take = min(need, self.n - self.pos)
            idx.append(self.perm[self.pos:self.pos + take])
            self.pos += take
            need -= take
        return np.concatenate(idx)


def _on_model_device(fn):
    """Run a model method with the model's GPU as the current CUDA device (streams, events and the staging copies below
    are taken from the current device)."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        dev = self.engine.device
        if torch.cuda.current_device() == dev.index:
            return fn(self, *args, **kwargs)
        with torch.cuda.device(dev):
            return fn(self, *args, **kwargs)
    return wrapper


class DSDGP(Parameterized):
    """Doubly Stochastic Deep Gaussian Process Model."""

    multitask = False

    def __init__(self, X, Y, Z, layers, likelihood, num_latent=None, minibatch_size=None, num_samples=1,
                 mean_function=Zero(), name=None, seed=1234, device=None):
        """
        - X is a data matrix, size N x D; Y is a data matrix, size N x R; Z inducing inputs, size M x D.
        - layers is an instance of Sequential; likelihood a Gaussian / SwitchedLikelihood.
        - minibatch_size, if not None, turns on minibatching; num_samples is the number of Monte Carlo samples.
        - seed keys the in-kernel Philox noise; device is the CUDA device used at compile() time.
        """
        super().__init__(name=name)
        X, Y, Z = (np.asarray(a, dtype=np.float64) for a in (X, Y, Z))
        assert X.shape[0] == Y.shape[0]
        assert Z.shape[1] == X.shape[1]
        self.num_data, D_X = X.shape
        self.num_samples = num_samples
        self.D_Y = num_latent or Y.shape[1]
        self.mean_function = mean_function      # kept for API parity; unused by the reference too
        layers.initialize_params(X, Z)
        if layers._initialized:
            self.layers = layers
        else:
            raise ValueError("Layers were not initialized")
        self.dims = self.layers.get_dims()
        self.likelihood = likelihood
        self.minibatch_size = minibatch_size
        self._mb = None if minibatch_size is None else _Minibatch(self.num_data, minibatch_size, seed=0)
        # data stays on the host (pinned when CUDA is present); every evaluation copies its minibatch to the device
        self.X = torch.from_numpy(X.copy())
        self.Y = torch.from_numpy(Y.copy())
        if torch.cuda.is_available():
            self.X, self.Y = self.X.pin_memory(), self.Y.pin_memory()
        self.seed = int(seed)
        self._step = 0
        self._device = device
        self._engine = None
        self.chunk_rows = None      # optional bound on minibatch rows processed per kernel pass (memory)
        self.predict_chunk_rows = 1 << 15   # test rows pushed through the cascade per pass in predict_* (x S samples)
        self._dist = None
        self._staging = None
        self._prefetched = None
        # Cholesky status words are read back every this many evaluations (0 = only on demand through
        # engine.check_status()).  The read is asynchronous: the copy into pinned memory is queued behind the evaluation and
        # looked at while a LATER evaluation is being queued -- the next one if the failed evaluation has finished by then,
        # otherwise the one after -- or at once by engine.poll_status(): a failed factorisation raises one or two calls
        # later and the host never waits for the device (the reference surfaces it as TF's InvalidArgumentError)
        self.check_status_every = 1
        # replay each evaluation of elbo_and_grad() as one CUDA graph (fixed minibatch shape; Philox noise only)
        self.use_cuda_graph = False
        # under distribute(): all-reduce each layer's gradients as soon as they are final (overlapping the backward pass of
        # the earlier layers) and capture those NCCL collectives inside the step's CUDA graph
        self.overlap_all_reduce = True
        self.graph_collectives = True
        self._graphs = {}

    # ------------------------------------------------------------------------------------------------ plumbing
    def compile(self, device=None):
        """Move parameters to the GPU and build the kernel descriptors (gpflow Model.compile())."""
        from . import _cabi
        _cabi.require_cuda()
        dev = torch.device(device or self._device or f"cuda:{torch.cuda.current_device()}")
        self.to(dev)
        self._engine = Engine(self.layers.layers, self.likelihood, multitask=self.multitask, device=dev)
        return self

    @property
    def engine(self):
        if self._engine is None:
            self.compile()
        return self._engine

    def distribute(self, process_group=None):
        """Shard every minibatch's rows over the ranks of a torch.distributed process group (one process per GPU);
        the flat gradient buffer is summed with one NCCL all-reduce per evaluation."""
        import torch.distributed as dist
        self._dist = (dist, process_group)
        return self

    def release_graphs(self):
        """Drop every captured CUDA graph of this model.  Graphs that captured NCCL collectives must be gone before
        torch.distributed.destroy_process_group() (NCCL waits for them when it tears the communicator down)."""
        self._graphs.clear()
        import gc
        gc.collect()
        torch.cuda.synchronize()

    def _reducer(self):
        """Per-layer bucketed all-reduce overlapping the backward pass (parallel.BucketReducer), or -- with
        `overlap_all_reduce` off -- one collective over the whole flat buffer after the last kernel."""
        dist, group = self._dist
        if self.overlap_all_reduce:
            return BucketReducer(dist, group)
        return lambda t: dist.all_reduce(t, group=group)

    def _next_batch(self):
        """Host view of the next minibatch (aligned X, Y rows)."""
        if self._mb is None:
            return self.X, self.Y
        idx = torch.from_numpy(self._mb.next_indices())
        return self.X[idx], self.Y[idx]

    def _stage_next(self):
        """Gather THIS rank's rows of the next minibatch into pinned staging buffers and start their host-to-device copy
        on a side stream (double-buffered), so the copy of step i+1 overlaps the kernels of step i.  Returns
        (X_dev, Y_dev, ready_event, row_offset, global_rows)."""
        dev = self.engine.device
        B = self.num_data if self._mb is None else self.minibatch_size
        lo, hi = 0, B
        if self._dist is not None:
            dist, group = self._dist
            lo, hi = shard_range(B, dist.get_rank(group), dist.get_world_size(group))
        if self._staging is None:
            mk = lambda cols: [(torch.empty(hi - lo, cols, dtype=torch.float64).pin_memory(),
                                torch.empty(hi - lo, cols, dtype=torch.float64, device=dev)) for _ in range(2)]
            self._staging = {"X": mk(self.X.shape[1]), "Y": mk(self.Y.shape[1]), "slot": 0,
                             "stream": torch.cuda.Stream(device=dev), "free": [None, None]}
        st = self._staging
        slot = st["slot"]
        st["slot"] ^= 1
        (xh, xd), (yh, yd) = st["X"][slot], st["Y"][slot]
        if st["free"][slot] is not None:
            st["free"][slot].synchronize()         # the kernels that read this slot two steps ago are done
        if self._mb is None:
            xh.copy_(self.X[lo:hi])
            yh.copy_(self.Y[lo:hi])
        else:
            idx = torch.from_numpy(self._mb.next_indices()[lo:hi])
            torch.index_select(self.X, 0, idx, out=xh)
            torch.index_select(self.Y, 0, idx, out=yh)
        with torch.cuda.stream(st["stream"]):
            xd.copy_(xh, non_blocking=True)
            yd.copy_(yh, non_blocking=True)
            ready = st["stream"].record_event()
        return xd, yd, ready, lo, B, slot

    def _eval_kwargs(self, X, Y, eps, seed):
        kw = dict(X=X, Y=Y, S=self.num_samples, eps=eps, seed=self.seed + self._step if seed is None else seed,
                  num_data=self.num_data, chunk_rows=self.chunk_rows)
        if self._dist is not None:
            dist, group = self._dist
            W, r = dist.get_world_size(group), dist.get_rank(group)
            B = X.shape[0]
            lo, hi = shard_range(B, r, W)
            kw.update(X=X[lo:hi], Y=Y[lo:hi], row_offset=lo, global_batch=B, world_size=W,
                      all_reduce=self._reducer())
            if eps is not None:
                kw["eps"] = [e[:, lo:hi] for e in eps]
        return kw

    # ------------------------------------------------------------------------------------------------ reference API
    def _propagate(self, Xnew, full_cov=False, num_samples=1, eps=None, seed=None):
        """Propagate points Xnew through the DGP cascade: (Fs, Fmeans, Fvars), lists of [S, N, .] CUDA tensors."""
        if full_cov:
            return self.engine.propagate_full_cov(Xnew, num_samples, z=eps, seed=self.seed if seed is None else seed)
        return self.engine.propagate(Xnew, num_samples, eps=eps, seed=self.seed if seed is None else seed,
                                     chunk_rows=self.predict_chunk_rows)

    def _build_predict(self, Xnew, full_cov=False, num_samples=1, eps=None, seed=None):
        if full_cov:
            _, Fmeans, Fvars = self._propagate(Xnew, True, num_samples, eps=eps, seed=seed)
            return Fmeans[-1], Fvars[-1]
        _, Fmeans, Fvars = self.engine.propagate(Xnew, num_samples, eps=eps, seed=self.seed if seed is None else seed,
                                                 chunk_rows=self.predict_chunk_rows, last_only=True)
        return Fmeans[-1], Fvars[-1]

    @_on_model_device
    def _build_likelihood(self, eps=None, seed=None, X=None, Y=None):
        """Variational bound on the marginal likelihood (dgplib/dsdgp.py:107-130) of the next minibatch, as a CUDA
        scalar attached to torch.autograd (call .backward() or hand it to a torch optimiser).  The node's forward runs
        the fused forward AND backward kernels (as one CUDA graph when use_cuda_graph is set)."""
        if X is not None:
            self._step += 1
            return elbo_autograd(self.engine, **self._eval_kwargs(X, Y, eps, seed))
        kw, finish = self._staged_step(eps, seed)
        out = elbo_autograd(self.engine, **kw)
        finish()
        return out

    def compute_log_likelihood(self, eps=None, seed=None):
        """gpflow Model.compute_log_likelihood(): the ELBO of one minibatch as a Python float."""
        with torch.no_grad():
            return float(self._build_likelihood(eps=eps, seed=seed))

    @_on_model_device
    def elbo_and_grad(self, X=None, Y=None, eps=None, seed=None):
        """One fused ELBO + gradient evaluation; returns (elbo tensor, grads of engine.params()).  With X=None the next
        minibatch comes from the model's own host data through the pinned, double-buffered staging path."""
        eng = self.engine
        if X is not None:
            self._step += 1
            with torch.no_grad():
                return eng.elbo(eng.native_params(), need_grad=True, **self._eval_kwargs(X, Y, eps, seed))
        kw, finish = self._staged_step(eps, seed)
        graph = kw.pop("_graph", None)
        check = kw.pop("check", False)
        if graph is not None:
            out = graph.run(None, kw["X"], kw["Y"], kw["seed"], all_reduce=kw.get("all_reduce"))
        else:
            with torch.no_grad():
                out = eng.elbo(eng.native_params(), need_grad=True, check=False, **kw)
        finish()
        if check:
            eng.post_status_read()
        return out

    def _staged_step(self, eps, seed):
        """Arguments of Engine.elbo for the next minibatch taken from the model's own host data through the pinned,
        double-buffered staging path, plus a callback to run once the step's kernels are enqueued."""
        eng = self.engine
        eng.poll_status(block=False)   # Cholesky status words of earlier evaluations that have arrived (posted asynchronously)
        staged = self._prefetched or self._stage_next()
        self._prefetched = None
        xd, yd, ready, lo, B, slot = staged
        self._step += 1
        main = torch.cuda.current_stream()
        main.wait_event(ready)
        kw = dict(X=xd, Y=yd, S=self.num_samples, eps=eps, seed=self.seed + self._step if seed is None else seed,
                  num_data=self.num_data, chunk_rows=self.chunk_rows, row_offset=lo, global_batch=B,
                  check=bool(self.check_status_every) and self._step % self.check_status_every == 0)
        if self._dist is not None:
            dist, group = self._dist
            kw.update(world_size=dist.get_world_size(group), all_reduce=self._reducer())
            if eps is not None:        # injected noise is given for the whole minibatch: this rank's rows of it
                kw["eps"] = [torch.as_tensor(e)[:, lo:lo + xd.shape[0]] for e in eps]
        if self.use_cuda_graph and eps is None and seed is None:
            kw["_graph"] = self._graph_for(xd, yd, kw)

        def finish():
            self._staging["free"][slot] = main.record_event()
            if self._mb is not None:
                self._prefetched = self._stage_next()      # next step's rows travel while this step computes
        return kw, finish

    def _graph_for(self, xd, yd, kw):
        from .engine import GraphedElbo
        eng = self.engine
        key = (tuple(xd.shape), tuple(yd.shape), kw["row_offset"], kw["global_batch"], kw.get("world_size", 1),
               self.num_samples, self.num_data, self.chunk_rows, self.seed, GraphedElbo.storage_key(eng))
        g = self._graphs.get(key)
        if g is None:
            # live parameters: the positive transforms are captured with the evaluation and read the model's own tensors
            captured = kw.get("all_reduce") if (self.graph_collectives and self.overlap_all_reduce) else None
            g = GraphedElbo(eng, None, xd, yd, self.num_samples, num_data=self.num_data,
                            row_offset=kw["row_offset"], global_batch=kw["global_batch"],
                            world_size=kw.get("world_size", 1), chunk_rows=self.chunk_rows, base_seed=self.seed,
                            all_reduce=captured)
            self._graphs[key] = g
        return g

    def predict_f(self, Xnew, num_samples, eps=None, seed=None):
        """Mean and variance of the final layer at Xnew: two [S, N, D_Y] arrays (dgplib/dsdgp.py:133-139)."""
        mean, var = self._build_predict(Xnew, False, num_samples, eps=eps, seed=seed)
        return mean.cpu().numpy(), var.cpu().numpy()

    def predict_all_layers(self, Xnew, num_samples, eps=None, seed=None):
        """(Fs, Fmeans, Fvars) for every layer (dgplib/dsdgp.py:152-158)."""
        out = self._propagate(Xnew, False, num_samples, eps=eps, seed=seed)
        return tuple([t.cpu().numpy() for t in lst] for lst in out)

    def predict_f_full_cov(self, Xnew, num_samples, eps=None, seed=None):
        """Mean [S,N,D_Y] and covariance [S,N,N,D_Y] of the final layer at Xnew (dgplib/dsdgp.py:142-148); samples are
        propagated with the full covariance across the N points (N <= 16384).  eps: per layer [S, R_l, N] draws."""
        mean, var = self._build_predict(Xnew, True, num_samples, eps=eps, seed=seed)
        return mean.cpu().numpy(), var.cpu().numpy()

    def predict_all_layers_full_cov(self, Xnew, num_samples, eps=None, seed=None):
        """(Fs, Fmeans, Fvars) for every layer with full covariances (dgplib/dsdgp.py:161-167)."""
        out = self._propagate(Xnew, True, num_samples, eps=eps, seed=seed)
        return tuple([t.cpu().numpy() for t in lst] for lst in out)

    def predict_f_samples(self, Xnew, num_samples, eps=None, V=None, seed=None):
        """Samples of the final layer's posterior at Xnew, [num_samples, N, D_Y] (dgplib/dsdgp.py:170-185): one
        full-covariance propagation, then mu + chol(var + jitter I) V per output.  V [D_Y, N, num_samples] injects the
        final draws."""
        mu, var = self._build_predict(Xnew, True, 1, eps=eps, seed=seed)
        mu, var = mu[0], var[0]                                      # [N, D_Y], [N, N, D_Y]
        N, DY = mu.shape
        eng = self.engine
        if V is None:
            g = torch.Generator(device=eng.device)
            g.manual_seed((self.seed if seed is None else seed) + 7919)
            V = torch.randn(DY, N, num_samples, dtype=torch.float64, device=eng.device, generator=g)
        out = eng.chol_sample(var.permute(2, 0, 1).contiguous(), V, mu.t().contiguous())    # [D_Y, N, num_samples]
        return out.permute(2, 1, 0).contiguous().cpu().numpy()

    def predict_y_moments(self, Xnew, num_samples=10, eps=None, seed=None):
        """Moment-matched predictive mean and variance of held-out data over `num_samples` propagated samples
        (what the notebooks plot): the equal-weight Gaussian mixture of predict_y's per-sample
        predictions, mean = E_s[mu_s], var = E_s[var_s + mu_s^2] - mean^2.  Returns two [N, D_Y] arrays."""
        mean, var = self._build_predict(Xnew, False, num_samples, eps=eps, seed=seed)
        lik = self.likelihood
        if hasattr(lik, "likelihood_list"):
            raise NotImplementedError("predict_y_moments is defined for a Gaussian likelihood")
        var = var + lik.variance.constrained().detach()
        m = mean.mean(0)
        v = (var + mean * mean).mean(0) - m * m
        return m.cpu().numpy(), v.cpu().numpy()

    def predict_y(self, Xnew, num_samples=1, eps=None, seed=None):
        """Mean and variance of held-out data at Xnew (dgplib/dsdgp.py:188-194).  The reference propagates a single
        sample (`_build_predict(Xnew)` defaults to num_samples=1); num_samples > 1 is a superset for sweeps."""
        mean, var = self._build_predict(Xnew, False, num_samples, eps=eps, seed=seed)
        lik = self.likelihood
        if hasattr(lik, "likelihood_list"):
            # gpflow SwitchedLikelihood.predict_mean_and_var concatenates every likelihood's prediction on axis 1
            vs = [var + l.variance.constrained().detach() for l in lik.likelihood_list]
            mean, var = torch.cat([mean] * len(vs), 1), torch.cat(vs, 1)
        else:
            var = var + lik.variance.constrained().detach()
        return mean.cpu().numpy(), var.cpu().numpy()
