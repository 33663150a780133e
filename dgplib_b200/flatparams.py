"""All parameters of a compiled model in ONE flat unconstrained fp64 device buffer, plus the index tables that let two
kernel launches replace the ~60 small torch launches that used to surround every evaluation:

* `dgp_transform_params` assembles the composite constrained arrays the kernels read -- per conditional theta [T, 2+D]
  (kernel variance, summed White variances, lengthscales broadcast when the kernel is not ARD) and the likelihood
  variances -- from the unconstrained entries through gpflow's Log1pe transform softplus(u) + 1e-6.  Z, q_mu, q_sqrt and a
  Linear mean have the identity transform: the kernels read them straight from the flat buffer.
* `dgp_adam_step` walks the flat buffer once: gradient w.r.t. the constrained arrays gathered from the evaluation's flat
  [elbo | gradients] buffer, softplus chain factor, tf.train.AdamOptimizer update (dgplib_b200/training.py).

The torch.nn.Parameters of the model keep working (they become views of the flat buffer), so `_build_likelihood()` as an
autograd node and any torch optimiser still see the same storage."""
import ctypes

import numpy as np
import torch

from . import _cabi, settings
from .mean_functions import Linear
from .params import Param

_F64 = torch.float64


def _params_of(module):
    return [m for m in module.modules() if isinstance(m, Param)]


class FlatParams:
    def __init__(self, engine):
        self.eng = engine
        dev = engine.device
        mods, seen = [], set()
        roots = list(engine.layers) + ([engine.likelihood] if engine.likelihood is not None else [])
        for root in roots:
            for p in _params_of(root):
                if id(p.raw) not in seen:
                    seen.add(id(p.raw))
                    mods.append(p)
        self.mods = mods
        sizes = [p.raw.numel() for p in mods]
        starts = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        self.n = int(starts[-1])
        self.start = {id(p.raw): int(s) for p, s in zip(mods, starts[:-1])}
        self.u = torch.empty(self.n, dtype=_F64, device=dev)
        flags = np.zeros(self.n, dtype=np.uint8)
        with torch.no_grad():
            for p, s, n in zip(mods, starts[:-1], sizes):
                view = self.u[s:s + n].view(p.raw.shape)
                view.copy_(p.raw.detach().to(device=dev, dtype=_F64))
                p.raw.data = view                      # the nn.Parameter is now a view of the flat buffer
                flags[s:s + n] = (1 if p.positive else 0) | (2 if p.raw.requires_grad else 0)
        self.key = self.storage_key(engine)
        self.native_version = 0        # bumped by every in-place update the library's own kernels make (adam_step)
        self._build_tables(flags)

    @staticmethod
    def storage_key(engine):
        """Identity of every parameter storage + its trainable flag: a re-assigned tensor or a set_trainable() call needs
        new tables (and a new capture of any CUDA graph that reads them)."""
        return tuple((p.data_ptr(), p.requires_grad) for p in engine.raw_parameters())

    # ------------------------------------------------------------------------------------------------ tables
    def _uidx(self, param):
        return self.start[id(param.raw)]

    def _build_tables(self, flags):
        eng, dev = self.eng, self.eng.device
        # ---- composite constrained arrays: [theta of every conditional | likelihood variances]
        crow, cidx, cpos = [0], [], []

        def term(param, k=0):
            cidx.append(self._uidx(param) + k)
            cpos.append(1 if param.positive else 0)

        self.theta_slices, comp_hyper = [], []      # comp_hyper: per composite element, the (param, k) terms
        for conds in eng.conds:
            row = []
            for c in conds:
                o0 = len(crow) - 1
                for (var, whites, ls) in c.kernel.task_param_refs():
                    term(var); crow.append(len(cidx)); comp_hyper.append([(var, 0)])
                    for w in whites:
                        term(w)
                    crow.append(len(cidx)); comp_hyper.append([(w, 0) for w in whites])
                    for k in range(c.D):
                        kk = k if ls.raw.numel() > 1 else 0
                        term(ls, kk); crow.append(len(cidx)); comp_hyper.append([(ls, kk)])
                row.append((o0, c.T, 2 + c.D))
            self.theta_slices.append(row)
        self.lik_slice = None
        lik = eng.likelihood
        if lik is not None:
            o0 = len(crow) - 1
            liks = list(lik.likelihood_list) if hasattr(lik, "likelihood_list") else [lik]
            for l in liks:
                term(l.variance); crow.append(len(cidx)); comp_hyper.append([(l.variance, 0)])
            self.lik_slice = (o0, len(liks))
        self.n_c = len(crow) - 1
        self.c = torch.zeros(max(self.n_c, 1), dtype=_F64, device=dev)
        t32 = lambda a: torch.tensor(np.asarray(a, dtype=np.int32), dtype=torch.int32, device=dev)
        self.crow, self.cidx = t32(crow), t32(cidx if cidx else [0])
        self.cpos = torch.tensor(np.asarray(cpos if cpos else [0], dtype=np.uint8), dtype=torch.uint8, device=dev)

        # ---- images of every u element in the flat [elbo | gradients] buffer of Engine.elbo (params() order)
        views = self.views()
        offs, total = eng.grad_layout(views)
        self.grad_total = total
        gidx = np.full(self.n, -1, dtype=np.int64)
        multi = {}

        def add(u0, g):
            """u elements u0 .. u0+len(g)-1 have one more image each, at the flat gradient indices g."""
            g = np.asarray(g, dtype=np.int64)
            idx = np.arange(u0, u0 + g.size)
            fresh = (gidx[idx] == -1)
            for j in idx[~fresh]:
                if j not in multi:
                    multi[j] = [int(gidx[j])]
                    gidx[j] = -2
            gidx[idx[fresh]] = g[fresh]
            for j, gi in zip(idx[~fresh], g[~fresh]):
                multi[int(j)].append(int(gi))

        it = iter(offs)
        comp_base = {}     # composite element index -> flat gradient index
        for layer, conds, th_row in zip(eng.layers, eng.conds, self.theta_slices):
            for c, (o0, T, w) in zip(conds, th_row):
                oZ, oT = next(it), next(it)
                Zp = c.feature.Z
                add(self._uidx(Zp), oZ + np.arange(Zp.raw.numel()))
                for e in range(T * w):
                    comp_base[o0 + e] = oT + e
            o_mu, o_sqrt = next(it), next(it)
            add(self._uidx(layer.q_mu), o_mu + np.arange(layer.q_mu.raw.numel()))
            add(self._uidx(layer.q_sqrt), o_sqrt + np.arange(layer.q_sqrt.raw.numel()))
            if eng.mean_trainable(layer):
                oA, ob = next(it), next(it)
                mf = layer.mean_function
                add(self._uidx(mf.A), oA + np.arange(mf.A.raw.numel()))
                add(self._uidx(mf.b), ob + np.arange(mf.b.raw.numel()))
        if self.lik_slice is not None:
            oL = next(it)
            for e in range(self.lik_slice[1]):
                comp_base[self.lik_slice[0] + e] = oL + e
        for ci, terms in enumerate(comp_hyper):
            for (param, k) in terms:
                add(self._uidx(param) + k, [comp_base[ci]])
        grow, gimg = [0], []
        for row, j in enumerate(sorted(multi)):
            gimg.extend(multi[j])
            grow.append(len(gimg))
            gidx[j] = -2 - row
        self.gidx = torch.tensor(gidx.astype(np.int32), dtype=torch.int32, device=dev)
        self.grow, self.gimg = t32(grow), t32(gimg if gimg else [0])
        self.flags = torch.tensor(flags, dtype=torch.uint8, device=dev)

    # ------------------------------------------------------------------------------------------------ use
    def views(self):
        """Constrained parameter tensors in Engine.params() order, as views of the flat / composite buffers (valid after
        transform())."""
        eng, out = self.eng, []
        for layer, conds, th_row in zip(eng.layers, eng.conds, self.theta_slices):
            for c, (o0, T, w) in zip(conds, th_row):
                out.append(c.feature.Z.raw.data)
                out.append(self.c[o0:o0 + T * w].view(T, w))
            out.append(layer.q_mu.raw.data)
            out.append(layer.q_sqrt.raw.data)
            if eng.mean_trainable(layer):
                out.append(layer.mean_function.A.raw.data)
                out.append(layer.mean_function.b.raw.data)
        if self.lik_slice is not None:
            out.append(self.c[self.lik_slice[0]:self.lik_slice[0] + self.lik_slice[1]])
        return out

    def transform(self):
        """Launch dgp_transform_params on the current stream."""
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        _cabi.check(_cabi.lib().dgp_transform_params(p(self.u), p(self.crow), p(self.cidx), p(self.cpos), self.n_c,
                                                     settings.positive_lower, p(self.c), _cabi.stream_ptr()),
                    "dgp_transform_params")

    def adam_step(self, state, grad_flat, lr, beta1, beta2, eps):
        """One tf.train.AdamOptimizer step on -ELBO from the evaluation's flat [elbo | gradients] buffer."""
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        assert grad_flat.numel() >= self.grad_total
        self.native_version += 1
        _cabi.check(_cabi.lib().dgp_adam_step(p(self.u), p(state["m"]), p(state["v"]), self.n, p(grad_flat), p(self.gidx),
                                              p(self.grow), p(self.gimg), p(self.flags), lr, beta1, beta2, eps,
                                              p(state["step"]), _cabi.stream_ptr()), "dgp_adam_step")

    def new_adam_state(self):
        dev = self.eng.device
        return {"m": torch.zeros(self.n, dtype=_F64, device=dev), "v": torch.zeros(self.n, dtype=_F64, device=dev),
                "step": torch.zeros(1, dtype=torch.int64, device=dev), "key": self.key}
