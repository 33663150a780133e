"""Training step around the hot path: the reference trains with
`gpflow.train.AdamOptimizer(lr).minimize(model, maxiter=n)` (tests/test_dsdgp.py:56-60, notebooks cell 16), i.e. Adam on the
unconstrained variables of -ELBO, one minibatch per step.  Here the ELBO + gradient of every step is one fused evaluation
(`DSDGP._build_likelihood`, a single autograd node), the softplus chain runs in torch, and the update is torch's fused
multi-tensor Adam."""
import torch


class AdamOptimizer:
    """gpflow.train.AdamOptimizer stand-in: `AdamOptimizer(0.01).minimize(model, maxiter=100)`."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.betas, self.eps = learning_rate, (beta1, beta2), epsilon
        self._opt = None
        self._model = None

    def _optimizer(self, model):
        if self._opt is None or self._model is not model:
            model.engine            # compile: parameters live on the GPU from here on
            params = model.trainable_parameters()
            self._opt = torch.optim.Adam(params, lr=self.lr, betas=self.betas, eps=self.eps, fused=True)
            self._model = model
        return self._opt

    def minimize(self, model, maxiter=1000, callback=None):
        """Run `maxiter` Adam steps on -ELBO; returns the list of per-step ELBO values (CUDA scalars, read lazily)."""
        opt = self._optimizer(model)
        history = []
        for it in range(int(maxiter)):
            opt.zero_grad(set_to_none=True)
            elbo = model._build_likelihood()
            (-elbo).backward()
            opt.step()
            history.append(elbo.detach())
            if callback is not None:
                callback(it, elbo)
        return history
