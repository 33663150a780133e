"""Training step around the hot path: the reference trains with
`gpflow.train.AdamOptimizer(lr).minimize(model, maxiter=n)` (tests/test_dsdgp.py:56-60, notebooks cell 16), i.e.
tf.train.AdamOptimizer on the unconstrained variables of -ELBO, one minibatch per step.

Here a step is a handful of launches of the extension and nothing else: `dgp_transform_params` (Log1pe transforms of the
hyper-parameters), the fused ELBO + gradient evaluation, and `dgp_adam_step` (softplus chain factor + TF's Adam update on
the flat unconstrained buffer, csrc/optim.cu).  With `model.use_cuda_graph` the whole step is captured once per minibatch
shape and replayed with one launch (`GraphedTrainStep`); under `model.distribute()` the flat gradient buffer is
all-reduced between the evaluation and the update, so every rank takes the same step."""
import ctypes

import torch

from . import _cabi


class GraphedTrainStep:
    """One Adam step on -ELBO as a CUDA graph (fixed minibatch shape, Philox noise).

    Static inputs: the minibatch (copied from the model's staged device buffers before each replay) and the per-step
    Philox seed offset (a stream-ordered device write in front of the replay).  Parameters and the optimiser state are
    updated in place by the replay.  The eager warm-up steps a capture needs are undone completely: parameters, both Adam
    moments and the step counter are snapshotted before and restored after, so a captured run takes exactly the steps an
    eager run takes, whatever state the optimiser was in when the graph was built."""

    def __init__(self, model, opt, X, Y, kw):
        with torch.cuda.device(model.engine.device):
            self._build(model, opt, X, Y, kw)

    def _build(self, model, opt, X, Y, kw):
        eng = model.engine
        dev = eng.device
        self.model, self.eng, self.opt = model, eng, opt
        self.Xs, self.Ys = X.clone(), Y.clone()
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.base_seed = int(model.seed)
        self.kw = dict(S=model.num_samples, eps=None, seed=self.base_seed, num_data=model.num_data,
                       chunk_rows=model.chunk_rows, row_offset=kw.get("row_offset", 0),
                       global_batch=kw.get("global_batch", X.shape[0]), world_size=kw.get("world_size", 1),
                       all_reduce=kw.get("all_reduce"), check=False)
        flat, st = eng.flatten(), opt._state_for(eng)
        saved = [flat.u.clone(), st["m"].clone(), st["v"].clone(), st["step"].clone()]
        cur = torch.cuda.current_stream()
        warm = torch.cuda.Stream(device=dev)
        warm.wait_stream(cur)
        with torch.cuda.stream(warm):
            for _ in range(2):
                self._step_body()
        cur.wait_stream(warm)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        eng._seed_dev = self.seed_dev
        try:
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self.elbo = self._step_body()
        finally:
            eng._seed_dev = None
        with torch.no_grad():                                   # undo the warm-up steps
            for dst, src in zip([flat.u, st["m"], st["v"], st["step"]], saved):
                dst.copy_(src)
        self._pins = ([c.ws for conds in eng.conds for c in conds], dict(eng._rows), eng._lik_ws, flat, st)

    def _step_body(self):
        eng = self.eng
        with torch.no_grad():
            val, _ = eng.elbo(eng.native_params(), self.Xs, self.Ys, need_grad=True, **self.kw)
            self.opt._apply(eng)
        return val

    def run(self, X, Y, seed):
        """One training step on the staged minibatch (X, Y device tensors) with the given full Philox seed; returns the
        step's ELBO (a view of a static buffer, overwritten by the next replay)."""
        with torch.cuda.device(self.eng.device):
            self.Xs.copy_(X, non_blocking=True)
            self.Ys.copy_(Y, non_blocking=True)
            _cabi.check(_cabi.lib().dgp_set_u64(ctypes.c_void_p(self.seed_dev.data_ptr()),
                                                (int(seed) - self.base_seed) & 0x7FFFFFFFFFFFFFFF, _cabi.stream_ptr()),
                        "dgp_set_u64")
            self.graph.replay()
            self.eng.flatten().native_version += 1      # the replayed Adam launch changed the parameters
        return self.elbo


class AdamOptimizer:
    """gpflow.train.AdamOptimizer stand-in: `AdamOptimizer(0.01).minimize(model, maxiter=100)`.  The update rule is
    tf.train.AdamOptimizer's (lr_t = lr sqrt(1 - beta2^t) / (1 - beta1^t); u -= lr_t m / (sqrt(v) + epsilon)), applied to
    the unconstrained variables by one kernel launch over the model's flat parameter buffer."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.betas, self.eps = float(learning_rate), (float(beta1), float(beta2)), float(epsilon)
        self._state = None
        self._graphs = {}

    def release_graphs(self):
        """Drop the captured training-step graphs (before torch.distributed.destroy_process_group(), see
        DSDGP.release_graphs)."""
        self._graphs.clear()

    def _state_for(self, eng):
        """Adam moments + step counter (device) for the engine's flat buffer; created on first use and whenever the flat
        buffer was rebuilt (a parameter re-assigned or its trainable flag changed)."""
        flat = eng.flatten()
        if self._state is None or self._state["key"] != flat.key:
            self._state = flat.new_adam_state()
            self._graphs = {}
        return self._state

    def _apply(self, eng):
        """The update from the evaluation that just ran (its flat [elbo | gradients] buffer, already all-reduced)."""
        flat = eng.flatten()
        flat.adam_step(self._state_for(eng), eng._last_flat, self.lr, self.betas[0], self.betas[1], self.eps)

    def _graphed_step(self, model):
        kw, finish = model._staged_step(None, None)
        kw.pop("_graph", None)
        X, Y = kw["X"], kw["Y"]
        eng = model.engine
        self._state_for(eng)
        key = (tuple(X.shape), tuple(Y.shape), kw.get("row_offset", 0), kw.get("global_batch"), kw.get("world_size", 1),
               model.num_samples, model.num_data, model.chunk_rows, model.seed, eng.flatten().key)
        g = self._graphs.get(key)
        if g is None:
            g = self._graphs[key] = GraphedTrainStep(model, self, X, Y, kw)
        elbo = g.run(X, Y, kw["seed"]).clone()
        finish()
        if kw.get("check"):
            eng.post_status_read()     # looked at when the next step is queued (DSDGP._staged_step) / at the end of minimize
        return elbo

    def _eager_step(self, model):
        kw, finish = model._staged_step(None, None)
        kw.pop("_graph", None)
        eng = model.engine
        check = kw.pop("check", False)
        with torch.no_grad():
            val, _ = eng.elbo(eng.native_params(), need_grad=True, check=False, **kw)
            self._apply(eng)
        val = val.clone()
        finish()
        if check:
            eng.post_status_read()
        return val

    def minimize(self, model, maxiter=1000, callback=None):
        """Run `maxiter` Adam steps on -ELBO; returns the list of per-step ELBO values (CUDA scalars, read lazily)."""
        eng = model.engine
        history = []
        with torch.cuda.device(eng.device):
            self._state_for(eng)
            # NCCL collectives are captured with the step when the process group allows it (model.graph_collectives)
            graphable = bool(model.use_cuda_graph) and (
                model._dist is None or (model.graph_collectives and model.overlap_all_reduce))
            for it in range(int(maxiter)):
                elbo = self._graphed_step(model) if graphable else self._eager_step(model)
                history.append(elbo)
                if callback is not None:
                    callback(it, elbo)
            eng.poll_status()
        return history
