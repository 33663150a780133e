"""Training step around the hot path: the reference trains with
`gpflow.train.AdamOptimizer(lr).minimize(model, maxiter=n)` (tests/test_dsdgp.py:56-60, notebooks cell 16), i.e. Adam on the
unconstrained variables of -ELBO, one minibatch per step.  Here the ELBO + gradient of every step is one fused evaluation
(`DSDGP._build_likelihood`, a single autograd node), the softplus chain runs in torch, and the update is torch's fused
multi-tensor Adam.  With `model.use_cuda_graph` on one GPU the WHOLE step -- positive transforms, the fused evaluation,
the chain rule back to the unconstrained variables and the Adam update -- is captured once as a CUDA graph and replayed
per minibatch (`GraphedTrainStep`), which removes the ~60 small launches around the evaluation from every step."""
import torch

from .engine import elbo_autograd


class GraphedTrainStep:
    """One Adam step on -ELBO as a CUDA graph (fixed minibatch shape, Philox noise, one GPU).

    Static inputs: the minibatch (copied from the model's staged device buffers before each replay) and the per-step
    Philox seed offset (added on the device).  Parameters and the optimiser state are updated in place by the replay.
    The three eager warm-up steps PyTorch needs before a capture are undone (parameters restored, Adam state zeroed), so a
    captured run takes exactly the steps an eager run takes."""

    def __init__(self, model, opt, X, Y):
        with torch.cuda.device(model.engine.device):
            self._build(model, opt, X, Y)

    def _build(self, model, opt, X, Y):
        eng = model.engine
        dev = eng.device
        self.model, self.eng, self.opt = model, eng, opt
        self.Xs, self.Ys = X.clone(), Y.clone()
        self.seed_host = torch.zeros(1, dtype=torch.int64).pin_memory()
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.base_seed = int(model.seed)
        self.key = (tuple(X.shape), tuple(Y.shape))
        params = [p for group in opt.param_groups for p in group["params"]]
        saved = [p.detach().clone() for p in params]
        cur = torch.cuda.current_stream()
        warm = torch.cuda.Stream(device=dev)
        warm.wait_stream(cur)
        with torch.cuda.stream(warm):
            for _ in range(3):
                self._step_body()
        cur.wait_stream(warm)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        eng._seed_dev = self.seed_dev
        try:
            with torch.cuda.graph(self.graph):
                self.seed_dev.copy_(self.seed_host, non_blocking=True)
                self.elbo = self._step_body()
        finally:
            eng._seed_dev = None
        with torch.no_grad():                                   # undo the warm-up steps
            torch._foreach_copy_(params, saved)
            for state in opt.state.values():
                for v in state.values():
                    if torch.is_tensor(v):
                        v.zero_()

    def _step_body(self):
        m = self.model
        self.opt.zero_grad(set_to_none=True)
        elbo = elbo_autograd(self.eng, X=self.Xs, Y=self.Ys, S=m.num_samples, eps=None, seed=self.base_seed,
                             num_data=m.num_data, chunk_rows=m.chunk_rows, row_offset=0, global_batch=self.Xs.shape[0],
                             check=False)
        (-elbo).backward()
        self.opt.step()
        return elbo.detach()

    def run(self, X, Y, seed):
        """One training step on the staged minibatch (X, Y device tensors) with the given full Philox seed; returns the
        step's ELBO (a view of a static buffer, overwritten by the next replay)."""
        with torch.cuda.device(self.eng.device):
            self.Xs.copy_(X, non_blocking=True)
            self.Ys.copy_(Y, non_blocking=True)
            self.seed_host[0] = (int(seed) - self.base_seed) & 0x7FFFFFFFFFFFFFFF     # added to base_seed on the device
            self.graph.replay()
        return self.elbo


class AdamOptimizer:
    """gpflow.train.AdamOptimizer stand-in: `AdamOptimizer(0.01).minimize(model, maxiter=100)`."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.betas, self.eps = learning_rate, (beta1, beta2), epsilon
        self._opt = None
        self._model = None
        self._graphs = {}

    def _graphable(self, model):
        return bool(model.use_cuda_graph) and model._dist is None

    def _optimizer(self, model):
        if self._opt is None or self._model is not model:
            model.engine            # compile: parameters live on the GPU from here on
            params = model.trainable_parameters()
            # capturable always: the step counter lives on the device, so model.use_cuda_graph may be switched on
            # between two minimize() calls without losing the optimiser state
            self._opt = torch.optim.Adam(params, lr=self.lr, betas=self.betas, eps=self.eps, fused=True, capturable=True)
            self._model = model
            self._graphs = {}
        return self._opt

    def _graphed_step(self, model, opt):
        kw, finish = model._staged_step(None, None)
        kw.pop("_graph", None)
        X, Y = kw["X"], kw["Y"]
        key = (tuple(X.shape), tuple(Y.shape))
        g = self._graphs.get(key)
        if g is None:
            g = self._graphs[key] = GraphedTrainStep(model, opt, X, Y)
        elbo = g.run(X, Y, kw["seed"]).clone()
        if kw.get("check"):
            model.engine.check_status()
        finish()
        return elbo

    def minimize(self, model, maxiter=1000, callback=None):
        """Run `maxiter` Adam steps on -ELBO; returns the list of per-step ELBO values (CUDA scalars, read lazily)."""
        opt = self._optimizer(model)
        history = []
        with torch.cuda.device(model.engine.device):
            self._loop(model, opt, int(maxiter), callback, history)
        return history

    def _loop(self, model, opt, maxiter, callback, history):
        for it in range(maxiter):
            if self._graphable(model):
                elbo = self._graphed_step(model, opt)
            else:
                opt.zero_grad(set_to_none=True)
                elbo = model._build_likelihood()
                (-elbo).backward()
                opt.step()
                elbo = elbo.detach()
            history.append(elbo)
            if callback is not None:
                callback(it, elbo)
