"""Host-side driver of the DSDGP hot path: packs the model's parameters into C-ABI descriptors, owns the device
workspaces and row buffers, and sequences the sm_100a kernels of libdgp_b200.so for

    propagate      DSDGP._propagate            dgplib/dsdgp.py:84-99   (+ multitask_dsdgp.py:9-26)
    elbo / grad    DSDGP._build_likelihood     dgplib/dsdgp.py:107-130 (+ tf.gradients of it)

PyTorch is used for device memory, streams and (multi-GPU) the NCCL all-reduce of the flat gradient buffer; every
arithmetic step is a kernel of the extension.  There is no CPU / eager fallback.
"""
import ctypes

import numpy as np
import torch

from . import _cabi, settings
from .mean_functions import Linear
from .parallel import kl_weight

_F64 = torch.float64


def _dev_tensor(x, device):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
    return x.detach().to(device=device, dtype=_F64).contiguous()


def _on_device(fn):
    """Run an Engine method with the engine's GPU as the current CUDA device: the C-ABI launches on the current device, so
    a model compiled for cuda:1 must not run its kernels under another device's context."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        if torch.cuda.current_device() == self.device.index:
            return fn(self, *args, **kwargs)
        with torch.cuda.device(self.device):
            return fn(self, *args, **kwargs)
    return wrapper


class _Cond:
    """One whitened conditional (kernel + inducing points + a column slice of q_mu / q_sqrt): descriptor + workspace."""

    def __init__(self, kernel, feature, col0, R, layer, Dx):
        self.kernel, self.feature, self.layer = kernel, feature, layer
        parts = kernel.task_kernels()          # raises NotImplementedError for unsupported kernels
        self.T = len(parts)
        self.kinds = [p[0] for p in parts]
        d = _cabi.CondDesc()
        d.M = layer.num_inducing
        d.Dx = Dx
        if kernel.switched:
            # SwitchedKernel strips the id column, then slices (dgplib/specialized_kernels.py:34-38)
            d.D = min(kernel.input_dim, Dx - 1)
        else:
            d.D = kernel.input_dim
        d.R, d.ldr, d.col0, d.T = R, layer.output_dim, col0, self.T
        for t, k in enumerate(self.kinds):
            d.kind[t] = k
        d.jitter = settings.jitter_level
        self.desc = d
        self.D = d.D
        self.ws = None
        self.ws_rows = -1
        self.ws_bwd = False
        self._keep = None

    def theta(self):
        """[T, 2 + D] constrained hyper-parameters (variance, White variance or 0, lengthscales), autograd-attached."""
        rows = []
        for (_, var, white, ls) in self.kernel.task_kernels():
            w = torch.zeros((), dtype=_F64, device=var.device) if white is None else white
            if ls.shape[0] != self.D:
                ls = ls[:self.D]
            rows.append(torch.cat([var.reshape(1), w.reshape(1), ls]))
        return torch.stack(rows)

    def bind(self, Z, theta, q_mu, q_sqrt, mean_A, mean_b):
        d = self.desc
        self._keep = (Z, theta, q_mu, q_sqrt, mean_A, mean_b)
        d.Z, d.theta, d.q_mu, d.q_sqrt = Z.data_ptr(), theta.data_ptr(), q_mu.data_ptr(), q_sqrt.data_ptr()
        d.mean_A = mean_A.data_ptr() if mean_A is not None else None
        d.mean_b = mean_b.data_ptr() if mean_b is not None else None

    def ensure_ws(self, n_rows, need_bwd, device):
        # the forward layout does not depend on n_rows (only the backward pass keeps row-sized buffers in the workspace):
        # any existing workspace serves a forward-only call, so a predict between two training steps never replaces the
        # backward-capable workspace a captured CUDA graph may point into
        if self.ws is not None and (not need_bwd or (self.ws_bwd and self.ws_rows >= n_rows)):
            return
        nbytes = _cabi.lib().dgp_cond_ws_bytes(ctypes.byref(self.desc), n_rows, 1 if need_bwd else 0)
        if nbytes == 0:
            raise _cabi.DgpError("dgp_cond_ws_bytes rejected the descriptor")
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.ws_rows, self.ws_bwd = n_rows, bool(need_bwd)

    def ref(self):
        return ctypes.byref(self.desc)


class Engine:
    """Compiled form of (Sequential, likelihood): the thing DSDGP calls every step."""

    def __init__(self, layers, likelihood=None, multitask=False, device=None):
        _cabi.require_cuda()
        _cabi.lib()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        # stream placement of the parameter stages, Gram reductions and parameter tails; bench.py's per-kernel timing pass
        # puts everything on the main stream so that an event pair times one kernel and not its wait for SMs
        self.gram_on_side_stream = True
        self.params_on_side_streams = True
        self._prep_key = None      # parameter version the forward halves of the workspaces were prepared for
        self.layers = list(layers)
        self.likelihood = likelihood
        self.multitask = bool(multitask)
        self.conds = []            # per layer: list of _Cond
        for layer in self.layers:
            Dx = layer.feature[0].Z.shape[1] if not hasattr(layer.feature, "Z") else layer.feature.Z.shape[1]
            self.conds.append([_Cond(k, f, c0, R, layer, Dx) for (k, f, c0, R) in layer.conditionals()])
        self._rows = {}
        self._lik_ws = None
        self._flat = None
        self.last_status = None
        self.timers = None         # set to {} to collect CUDA-event timings per kernel class (bench.py)

    def _timed(self, key, fn):
        """Run fn() (which enqueues kernels on the current stream); when timers are on, bracket it with CUDA events
        recorded on that same stream."""
        if self.timers is None:
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        self.timers.setdefault(key, []).append((e0, e1))
        return out

    def timer_summary(self):
        """{key: (launches, total ms)} -- call after a stream sync."""
        out = {}
        for k, evs in (self.timers or {}).items():
            out[k] = (len(evs), sum(a.elapsed_time(b) for a, b in evs))
        return out

    # ------------------------------------------------------------------------------------------ parameters
    def _lik_variances(self):
        lik = self.likelihood
        if hasattr(lik, "likelihood_list"):
            return torch.stack([l.variance.constrained() for l in lik.likelihood_list])
        return lik.variance.constrained().reshape(1)

    def raw_parameters(self):
        """The unconstrained torch Parameters behind params() (every Param of the layers and the likelihood)."""
        from . import params as _params
        epoch = _params.structure_epoch()
        cached = self.__dict__.get("_raw_params")
        if cached is None or cached[0] != epoch:
            # the walk over the module tree costs ~10 us per layer and used to run a dozen times per evaluation: kept until
            # the set of Parameters can have changed (params.structure_epoch)
            out = []
            for layer in self.layers:
                out.extend(layer.parameters())
            if self.likelihood is not None:
                out.extend(self.likelihood.parameters())
            cached = self._raw_params = (epoch, out)
        return cached[1]

    @staticmethod
    def mean_trainable(layer):
        """A Linear mean whose A or b is trainable: only an OutputLayer built with its own Linear mean (the reference
        freezes the mean of Input / Hidden layers in initialize_forward, dgplib/layers.py:184-186, :202-204, but not of an
        OutputLayer, :211-218)."""
        mf = layer.mean_function
        return isinstance(mf, Linear) and (mf.A.trainable or mf.b.trainable)

    def params(self):
        """Constrained, autograd-attached parameter tensors in the flat order used by the gradient buffer:
        per layer [Z_c, theta_c for each conditional c] + [q_mu, q_sqrt] (+ [mean A, mean b] of a trainable Linear mean);
        then the likelihood variances."""
        out = []
        for layer, conds in zip(self.layers, self.conds):
            for c in conds:
                out.append(c.feature.Z.constrained())
                out.append(c.theta())
            out.append(layer.q_mu.constrained())
            out.append(layer.q_sqrt.constrained())
            if self.mean_trainable(layer):
                out.append(layer.mean_function.A.constrained())
                out.append(layer.mean_function.b.constrained())
        if self.likelihood is not None:
            out.append(self._lik_variances())
        return out

    def flatten(self):
        """The model's parameters as one flat unconstrained buffer + transform / chain-rule tables (flatparams.py);
        rebuilt when a parameter tensor was re-assigned or its trainable flag changed."""
        from .flatparams import FlatParams
        if self._flat is None or self._flat.key != FlatParams.storage_key(self):
            with torch.cuda.device(self.device):
                self._flat = FlatParams(self)
        return self._flat

    def native_params(self):
        """params() without torch arithmetic: one dgp_transform_params launch on the current stream, then views of the
        flat buffers (detached; what the non-autograd entry points and the captured graphs use)."""
        flat = self.flatten()
        flat.transform()
        return flat.views()

    def _bind(self, params):
        """params: detached contiguous CUDA tensors in params() order."""
        it = iter(params)
        for layer, conds in zip(self.layers, self.conds):
            zt = [(next(it), next(it)) for _ in conds]
            q_mu, q_sqrt = next(it), next(it)
            mA = mb = None
            if self.mean_trainable(layer):
                mA, mb = next(it), next(it)
            elif isinstance(layer.mean_function, Linear):
                mA = _dev_tensor(layer.mean_function.A.constrained(), self.device)
                mb = _dev_tensor(layer.mean_function.b.constrained(), self.device)
                if mA.shape != (conds[0].desc.Dx, layer.output_dim):
                    raise ValueError(f"Linear mean A has shape {tuple(mA.shape)}, expected "
                                     f"{(conds[0].desc.Dx, layer.output_dim)}")
            for c, (Z, th) in zip(conds, zt):
                c.bind(Z, th, q_mu, q_sqrt, mA, mb)
        self._lik_var = next(it) if self.likelihood is not None else None

    def _detached(self, params=None):
        params = self.params() if params is None else params
        return [_dev_tensor(p, self.device) for p in params]

    # ------------------------------------------------------------------------------------------ buffers
    def _rows_of(self, l, S, nb):
        """Rows the conditionals of layer l see for nb data rows: the first layer is evaluated once per data row
        (tile_over_samples makes its S copies identical), every later layer on all S * nb samples."""
        return nb if l == 0 else S * nb

    def _row_buffers(self, S, nb, save):
        key = (S, nb, bool(save))
        buf = self._rows.get(key)
        if buf is None:
            dev = self.device
            n = S * nb
            buf = {"mean": [], "var": [], "F": [], "A": [], "U": [], "K": [], "gX": []}
            for l, (layer, conds) in enumerate(zip(self.layers, self.conds)):
                R = layer.output_dim
                ldf = R + (1 if self.multitask else 0)
                rows = self._rows_of(l, S, nb)
                buf["mean"].append(torch.empty(rows, R, dtype=_F64, device=dev))
                buf["var"].append(torch.empty(rows, R, dtype=_F64, device=dev))
                buf["F"].append(torch.empty(n, ldf, dtype=_F64, device=dev))
                if save:
                    buf["A"].append([torch.empty(rows, _cabi.mpad(layer.num_inducing), dtype=_F64, device=dev)
                                     for _ in conds])
                    # U_r = tril(q_sqrt_r)^T A of every output, kept for the backward pass (the reference's LTA tensor)
                    buf["U"].append([torch.empty(_cabi.lib().dgp_usave_bytes(c.ref(), rows) // 8, dtype=_F64, device=dev)
                                     for c in conds])
                    # Kuf of an RBF conditional (or a SwitchedKernel of RBFs): its backward pass then needs no exp
                    # (k' = -k / 2)
                    buf["K"].append([torch.empty(rows, _cabi.mpad(layer.num_inducing), dtype=_F64, device=dev)
                                     if all(k == _cabi.DGP_KIND_RBF for k in c.kinds[:c.T]) else None for c in conds])
                    buf["gX"].append(torch.empty(n, conds[0].desc.Dx, dtype=_F64, device=dev) if l > 0 else None)
            if save:
                RL, R0 = self.layers[-1].output_dim, self.layers[0].output_dim
                rows_last = self._rows_of(len(self.layers) - 1, S, nb)
                buf["g_mu"] = torch.empty(rows_last, RL, dtype=_F64, device=dev)
                buf["g_var"] = torch.empty(rows_last, RL, dtype=_F64, device=dev)
                buf["g_mean0"] = torch.empty(nb, R0, dtype=_F64, device=dev)
                buf["g_var0"] = torch.empty(nb, R0, dtype=_F64, device=dev)
            if len(self._rows) > 8:
                self._rows.clear()
            self._rows[key] = buf
        return buf

    @staticmethod
    def _groupable(conds):
        """A MultikernelLayer whose conditionals can share one launch: 2..8 members of identical shape."""
        if not (2 <= len(conds) <= 8):
            return False
        d0 = conds[0].desc
        return all((c.desc.M, c.desc.D, c.desc.Dx, c.desc.R, c.desc.T, c.desc.ldr) ==
                   (d0.M, d0.D, d0.Dx, d0.R, d0.T, d0.ldr) for c in conds)

    def _side_streams(self):
        """One side stream per conditional (parameter stages of different conditionals are independent)."""
        if getattr(self, "_streams", None) is None:
            self._streams = [[torch.cuda.Stream(device=self.device) for _ in conds] for conds in self.conds]
        return self._streams

    def _comm_stream(self):
        if getattr(self, "_comm", None) is None:
            self._comm = torch.cuda.Stream(device=self.device)
        return self._comm

    def _prepare(self, S, nb, need_bwd, kl):
        """Parameter-only stage of every conditional (Kuu, Cholesky, inverse factor, tiled operands, KL).  The
        conditionals are independent of each other and of the data, so each chain runs on its own stream; returns, per
        layer, the events the main stream waits on just before that layer's first row kernel."""
        lib = _cabi.lib()
        main = torch.cuda.current_stream()
        start = main.record_event()
        events, i = [], 0
        for l, (conds, sides) in enumerate(zip(self.conds, self._side_streams())):
            evs = []
            for c, side in zip(conds, sides):
                if not self.params_on_side_streams:
                    side = main
                else:
                    side.wait_event(start)
                with torch.cuda.stream(side):
                    c.ensure_ws(self._rows_of(l, S, nb), need_bwd, self.device)
                    self._timed("prepare", lambda: _cabi.check(lib.dgp_cond_prepare(
                        c.ref(), _cabi.ptr(c.ws), c.ws.numel(), 1 if need_bwd else 0,
                        ctypes.c_void_p(kl[i:].data_ptr()) if kl is not None else None, _cabi.stream_ptr()),
                        "dgp_cond_prepare"))
                    if side is not main:
                        evs.append(side.record_event())
                i += 1
            events.append(evs)
        return events

    def _param_version(self):
        """Changes whenever a parameter of the model may have changed: torch bumps a parameter's version counter on every
        in-place write to it (Param.assign, optimisers, user code; the counters are per parameter -- `p.data = view` does
        not share the flat buffer's), the library's own Adam launch bumps `native_version`."""
        flat = self.flatten()
        return (flat.key, sum(p._version for p in self.raw_parameters()), flat.native_version)

    @_on_device
    def check_status(self):
        """Raises if a Cholesky factorisation failed (the reference surfaces this as TF's InvalidArgumentError)."""
        lib, st = _cabi.lib(), _cabi.stream_ptr()
        for l, conds in enumerate(self.conds):
            for c in conds:
                s = ctypes.c_int32(0)
                _cabi.check(lib.dgp_cond_status(_cabi.ptr(c.ws), st, ctypes.byref(s)), "dgp_cond_status")
                if s.value != 0:
                    raise _cabi.DgpError(f"Cholesky of Kuu failed in layer {l} at pivot {s.value - 1}: "
                                         "Kuu + jitter is not positive definite")

    @_on_device
    def post_status_read(self):
        """Queue the copy of every conditional's Cholesky status word into pinned host memory behind the work already on
        the current stream, without waiting for it; `poll_status()` looks at the result once that work has finished."""
        lib, st = _cabi.lib(), _cabi.stream_ptr()
        n = sum(len(c) for c in self.conds)
        pend = self.__dict__.setdefault("_status_pending", [])
        pool = self.__dict__.setdefault("_status_pool", [])
        if len(pend) >= 3:                      # never more than three reads in flight (three pinned buffers)
            self._check_status_entry(pend.pop(0), wait=True)
        host = pool.pop() if pool and pool[-1].numel() == n else torch.zeros(n, dtype=torch.int32).pin_memory()
        i = 0
        for conds in self.conds:
            for c in conds:
                _cabi.check(lib.dgp_cond_status_async(_cabi.ptr(c.ws), st, ctypes.c_void_p(host.data_ptr() + 4 * i)),
                            "dgp_cond_status_async")
                i += 1
        pend.append((torch.cuda.current_stream().record_event(), host))

    def _check_status_entry(self, entry, wait):
        ev, host = entry
        if wait:
            ev.synchronize()
        bad = host.nonzero()
        info = (int(bad[0]), int(host[int(bad[0])]) - 1) if bad.numel() else None
        self._status_pool.append(host)
        if info is not None:
            raise _cabi.DgpError(f"Cholesky of Kuu failed in conditional {info[0]} at pivot {info[1]}: "
                                 "Kuu + jitter is not positive definite")

    def poll_status(self, block=True):
        """Raises if an evaluation whose status was posted by post_status_read() had a failed Cholesky factorisation.
        block=True waits for every posted read; block=False (what DSDGP does while queueing the next evaluation) looks only
        at the reads that have already arrived, so the host never waits for the evaluation that is still running -- the
        device always has the next step queued behind the current one."""
        pend = getattr(self, "_status_pending", None)
        while pend:
            if not block and not pend[0][0].query():
                break
            self._check_status_entry(pend.pop(0), wait=True)

    # ------------------------------------------------------------------------------------------ forward
    def _seed_dev_ptr(self):
        sd = getattr(self, "_seed_dev", None)
        return ctypes.c_void_p(sd.data_ptr()) if sd is not None else None

    def _layer_seed(self, seed, l):
        return ctypes.c_uint64((seed + 0x9E3779B97F4A7C15 * (l + 1)) & 0xFFFFFFFFFFFFFFFF)

    def _forward_rows(self, X, S, eps, seed, row_offset, sample_stride, buf, save, last_F, prep_events):
        """One pass of the cascade over the data rows X [nb, Dx0] and their S samples.
        eps: list of [S*nb, R_l] tensors or None (Philox)."""
        lib, st = _cabi.lib(), _cabi.stream_ptr()
        main = torch.cuda.current_stream()
        nb = X.shape[0]
        n = S * nb
        inp, x_rows = X, nb
        L = len(self.layers)
        stride = sample_stride if sample_stride else nb
        for l, (layer, conds) in enumerate(zip(self.layers, self.conds)):
            for ev in prep_events[l]:
                main.wait_event(ev)
            want_F = last_F or l < L - 1
            Fbuf = buf["F"][l] if want_F else None
            rows = self._rows_of(l, S, nb)
            fused_F = want_F and l > 0        # layer 0 samples are drawn by dgp_sample_expand (S per data row)
            eps_ptr = _cabi.ptr(eps[l]) if (eps is not None and fused_F) else None
            F_ptr = _cabi.ptr(Fbuf) if fused_F else None
            ldf = Fbuf.shape[1] if fused_F else layer.output_dim
            grouped = False
            if self._groupable(conds):
                # a MultikernelLayer's conditionals in ONE launch over (conditional, row tile) pairs
                n_c = len(conds)
                descs = (ctypes.POINTER(_cabi.CondDesc) * n_c)(*[ctypes.pointer(c.desc) for c in conds])
                wss = (ctypes.c_void_p * n_c)(*[c.ws.data_ptr() for c in conds])
                saves = [(ctypes.c_void_p * n_c)(*[(b.data_ptr() if b is not None else None) for b in buf[k][l]])
                         if save else None for k in ("A", "U", "K")]
                rc = self._timed(f"fwd.l{l}", lambda: lib.dgp_cond_forward_group(
                    n_c, descs, wss, _cabi.ptr(inp), x_rows, rows, eps_ptr, self._layer_seed(seed, l), self._seed_dev_ptr(),
                    row_offset, nb, stride, _cabi.ptr(buf["mean"][l]), _cabi.ptr(buf["var"][l]), F_ptr, ldf,
                    saves[0], saves[1], saves[2], st))
                if rc != _cabi.DGP_ERR_UNSUPPORTED:
                    _cabi.check(rc, "dgp_cond_forward_group")
                    grouped = True
            for ci, c in enumerate(conds if not grouped else []):
                self._timed(f"fwd.l{l}", lambda: _cabi.check(lib.dgp_cond_forward(
                    c.ref(), _cabi.ptr(c.ws), _cabi.ptr(inp), x_rows, rows, eps_ptr,
                    self._layer_seed(seed, l), self._seed_dev_ptr(), row_offset, nb, stride,
                    _cabi.ptr(buf["mean"][l]), _cabi.ptr(buf["var"][l]), F_ptr, ldf,
                    _cabi.ptr(buf["A"][l][ci]) if save else None, _cabi.ptr(buf["U"][l][ci]) if save else None,
                    _cabi.ptr(buf["K"][l][ci]) if save else None, st), "dgp_cond_forward"))
            if l == 0 and want_F:
                self._timed("sample.l0", lambda: _cabi.check(lib.dgp_sample_expand(
                    _cabi.ptr(buf["mean"][0]), _cabi.ptr(buf["var"][0]), _cabi.ptr(eps[0]) if eps is not None else None,
                    self._layer_seed(seed, 0), self._seed_dev_ptr(), row_offset, stride, S, nb, layer.output_dim, layer.output_dim,
                    _cabi.ptr(Fbuf), Fbuf.shape[1], _cabi.ptr(X), X.shape[1], st), "dgp_sample_expand"))
            inp, x_rows = Fbuf, n

    @_on_device
    def propagate(self, X, S, eps=None, seed=0, params=None, chunk_rows=None, last_only=False):
        """Fs, Fmeans, Fvars as lists of [S, B, .] CUDA tensors (DSDGP._propagate, dgplib/dsdgp.py:84-99).
        chunk_rows bounds the data rows pushed through the cascade per pass (memory: S * chunk_rows rows are live per
        layer), last_only returns only the final layer's (mean, var) -- what predict_f / predict_y need."""
        with torch.no_grad():
            X = _dev_tensor(X, self.device)
            B = X.shape[0]
            Bc = B if not chunk_rows else max(1, min(B, int(chunk_rows)))
            if params is None:
                # the model's own parameters: the parameter stage (33 ms at M = 1024) is redone only when they changed
                # since the workspaces were last prepared -- a prediction sweep pays for it once
                self._bind(self.native_params())
                key = self._param_version()
                if key == self._prep_key and all(c.ws is not None for conds in self.conds for c in conds):
                    events = [[] for _ in self.conds]
                else:
                    events = self._prepare(S, Bc, False, None)
                    self._prep_key = key
            else:
                self._bind(self._detached(params))
                self._prep_key = None
                events = self._prepare(S, Bc, False, None)
            eps_full = None
            if eps is not None:
                eps_full = [_dev_tensor(e, self.device).reshape(S, B, -1) for e in eps]
            L = len(self.layers)
            dev = self.device
            keep = [L - 1] if last_only else list(range(L))
            out = {k: {l: None for l in keep} for k in ("F", "mean", "var")}
            for l in keep:
                R = self.layers[l].output_dim
                ldf = R + (1 if self.multitask else 0)
                out["mean"][l] = torch.empty(S, B, R, dtype=_F64, device=dev)
                out["var"][l] = torch.empty(S, B, R, dtype=_F64, device=dev)
                if not last_only:
                    out["F"][l] = torch.empty(S, B, ldf, dtype=_F64, device=dev)
            for b0 in range(0, B, Bc):
                b1 = min(B, b0 + Bc)
                nb = b1 - b0
                buf = self._row_buffers(S, nb, False)
                eps_c = None
                if eps_full is not None:
                    eps_c = [e[:, b0:b1].reshape(S * nb, -1).contiguous() for e in eps_full]
                self._forward_rows(X[b0:b1], S, eps_c, seed, b0, B, buf, False, not last_only, events)
                for l in keep:
                    for key in ("mean", "var"):
                        t = buf[key][l]
                        src = t.unsqueeze(0).expand(S, *t.shape) if (l == 0 and S > 1) else t.reshape(S, nb, -1)
                        out[key][l][:, b0:b1].copy_(src)
                    if not last_only:
                        out["F"][l][:, b0:b1].copy_(buf["F"][l].reshape(S, nb, -1))
            self.check_status()
            if last_only:
                return None, [out["mean"][L - 1]], [out["var"][L - 1]]
            return ([out["F"][l] for l in keep], [out["mean"][l] for l in keep], [out["var"][l] for l in keep])

    # ------------------------------------------------------------------------------------------ full covariance
    def _layer_full_cov(self, l, X, lib, st):
        """mean [N, R] and covariance [R, N, N] of layer l at the rows X [N, Dx] (one sample): the row kernel gives the
        mean and A = Lm^-1 Kuf, dgp_cond_fullcov turns A into Knn - A^T A + (Lq^T A)^T (Lq^T A)."""
        layer, conds = self.layers[l], self.conds[l]
        N, R = X.shape[0], layer.output_dim
        dev = self.device
        mean = torch.empty(N, R, dtype=_F64, device=dev)
        var_diag = torch.empty(N, R, dtype=_F64, device=dev)
        fvar = torch.empty(R, N, N, dtype=_F64, device=dev)
        for c in conds:
            A = torch.empty(N, _cabi.mpad(layer.num_inducing), dtype=_F64, device=dev)
            _cabi.check(lib.dgp_cond_forward(c.ref(), _cabi.ptr(c.ws), _cabi.ptr(X), N, N, None, 0, None, 0, 0, 0,
                                             _cabi.ptr(mean), _cabi.ptr(var_diag), None, R, _cabi.ptr(A), None, None, st),
                        "dgp_cond_forward")
            nbytes = lib.dgp_cond_fullcov_ws_bytes(c.ref(), N)
            if nbytes == 0:
                raise _cabi.DgpError("full-covariance prediction supports at most 16384 points per call")
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            out = fvar[c.desc.col0:c.desc.col0 + c.desc.R]
            _cabi.check(lib.dgp_cond_fullcov(c.ref(), _cabi.ptr(c.ws), _cabi.ptr(X), N, _cabi.ptr(A),
                                             ctypes.c_void_p(out.data_ptr()), _cabi.ptr(scratch), nbytes, st),
                        "dgp_cond_fullcov")
        return mean, fvar

    @_on_device
    def chol_sample(self, var, z, mean=None, jitter=None):
        """mean + chol(var + jitter I) z for a batch: var [Bt, N, N], z [Bt, N, nz], mean [Bt, N] or None -> [Bt, N, nz]
        (normal_sample's full-covariance branch, dgplib/utils.py:28-36)."""
        lib, st = _cabi.lib(), _cabi.stream_ptr()
        var, z = _dev_tensor(var, self.device), _dev_tensor(z, self.device)
        Bt, N, nz = z.shape
        mean = None if mean is None else _dev_tensor(mean, self.device)
        out = torch.empty(Bt, N, nz, dtype=_F64, device=self.device)
        nbytes = lib.dgp_chol_sample_ws_bytes(N, Bt)
        if nbytes == 0:
            raise _cabi.DgpError("full-covariance sampling supports at most 16384 points per call (128 covariance matrices per call above 1024)")
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        status = torch.zeros(Bt, dtype=torch.int32, device=self.device)
        _cabi.check(lib.dgp_chol_sample(_cabi.ptr(var), N * N, N, 1, N, Bt,
                                        settings.jitter_level if jitter is None else jitter, _cabi.ptr(z), nz,
                                        _cabi.ptr(mean) if mean is not None else None, N, 1, _cabi.ptr(out),
                                        _cabi.ptr(scratch), nbytes, ctypes.c_void_p(status.data_ptr()), st),
                    "dgp_chol_sample")
        bad = int(status.max())
        if bad:
            raise _cabi.DgpError(f"Cholesky of a predictive covariance failed at pivot {bad - 1}")
        return out

    @_on_device
    def propagate_full_cov(self, X, S, z=None, seed=0, params=None):
        """DSDGP._propagate(full_cov=True) (dgplib/dsdgp.py:84-99): Fs [S,N,R(+1)], Fmeans [S,N,R], Fvars [S,N,N,R] per
        layer.  z[l] [S, R_l, N] are the N(0,1) draws of normal_sample's full-covariance branch (Philox when None)."""
        lib, st = _cabi.lib(), _cabi.stream_ptr()
        with torch.no_grad():
            dev = self.device
            X = _dev_tensor(X, dev)
            self._bind(self._detached(params))
            self._prep_key = None
            N = X.shape[0]
            for evs in self._prepare(1, N, False, None):
                for ev in evs:
                    torch.cuda.current_stream().wait_event(ev)
            Fs, Ms, Vs = [[] for _ in self.layers], [[] for _ in self.layers], [[] for _ in self.layers]
            for s in range(S):
                inp = X
                for l, layer in enumerate(self.layers):
                    R = layer.output_dim
                    mean, fvar = self._layer_full_cov(l, inp, lib, st)
                    if z is not None:
                        zz = _dev_tensor(z[l], dev)[s].reshape(R, N, 1)
                    else:
                        zz = torch.empty(N, R, dtype=_F64, device=dev)
                        _cabi.check(lib.dgp_philox_normal(_cabi.ptr(zz), N, R, R, self._layer_seed(seed, l), s * N, 0, 0,
                                                          st), "dgp_philox_normal")
                        zz = zz.t().contiguous().reshape(R, N, 1)
                    F = self.chol_sample(fvar, zz, mean.t().contiguous())[:, :, 0].t().contiguous()   # [N, R]
                    if self.multitask:
                        F = torch.cat([F, inp[:, -1:]], 1)
                    Fs[l].append(F)
                    Ms[l].append(mean)
                    Vs[l].append(fvar.permute(2, 1, 0).contiguous())        # [R,N,N] -> [N,N,R] (dgplib/layers.py:77)
                    inp = F
            self.check_status()
            return ([torch.stack(f) for f in Fs], [torch.stack(m) for m in Ms], [torch.stack(v) for v in Vs])

    # ------------------------------------------------------------------------------------------ ELBO (+ grad)
    def grad_layout(self, params):
        """Offsets of [elbo, grads of params()...] inside the flat fp64 gradient buffer."""
        offs, o = [], 1
        for p in params:
            offs.append(o)
            o += p.numel()
        return offs, o

    @_on_device
    def elbo(self, params, X, Y, S, eps=None, seed=0, num_data=None, need_grad=True, row_offset=0,
             global_batch=None, world_size=1, all_reduce=None, chunk_rows=None, check=True):
        """ELBO (dgplib/dsdgp.py:107-130) of the minibatch rows X [B, Dx], Y [B, ldy] held by THIS rank, and -- if
        need_grad -- its gradient w.r.t. every tensor of params().  Returns (elbo 0-dim tensor, [grad tensors]).

        Multi-GPU: each rank passes its own row shard; `global_batch` is the whole minibatch size, `row_offset` the
        shard's first row in it, and `all_reduce(flat)` sums the flat buffer over ranks (NCCL).  The KL term enters
        with weight 1/world_size on every rank so the sum is exact."""
        lib, st = _cabi.lib(), _cabi.stream_ptr()
        with torch.no_grad():
            dev = self.device
            main = torch.cuda.current_stream()
            X, Y = _dev_tensor(X, dev), _dev_tensor(Y, dev)
            params = [_dev_tensor(p, dev) for p in params]
            self._bind(params)
            B = X.shape[0]
            Bg = global_batch if global_batch is not None else B
            num_data = Bg if num_data is None else num_data
            L = len(self.layers)
            # (num_data / B) / S (dgplib/dsdgp.py:121, :128-129); a 1-layer model's S samples are identical copies
            S_lik = S if L > 1 else 1
            weight = (float(num_data) / float(Bg)) / float(S_lik)
            ncond = sum(len(c) for c in self.conds)
            offs, total = self.grad_layout(params)
            flat = torch.empty(total + ncond, dtype=_F64, device=dev)   # [elbo | grads | kl per conditional]
            _cabi.check(lib.dgp_zero(ctypes.c_void_p(flat.data_ptr()), flat.numel() * 8, st), "dgp_zero")
            kl = flat[total:]
            Bc = B if not chunk_rows else max(1, min(B, int(chunk_rows)))
            self._prep_key = None       # prepared for whatever `params` holds, which this method cannot identify
            events = self._prepare(S, Bc, need_grad, kl)
            if self._lik_ws is None:
                self._lik_ws = torch.empty(lib.dgp_lik_ws_bytes(0, 1), dtype=torch.uint8, device=dev)
            RL = self.layers[-1].output_dim
            T_lik = self._lik_var.numel()
            ldy = Y.shape[1]
            g_lik = flat[offs[-1]:offs[-1] + T_lik]
            klw = kl_weight(world_size)
            eps_full = None
            if eps is not None:
                eps_full = [_dev_tensor(e, dev).reshape(S, B, -1) for e in eps]
            tail_events = []
            bucketed = need_grad and world_size > 1 and all_reduce is not None and hasattr(all_reduce, "bucket")

            def reduce_bucket(l, events):
                """All-reduce of layer l's gradient slots as soon as they are final: [layer l] for a middle layer, the last
                layer's together with the likelihood gradient behind it, the first layer's together with the ELBO in front
                of it.  Issued on the communication stream behind the events of the tails that wrote the slots."""
                lo = 0 if l == 0 else self._zt_offsets[l][0][0][1]
                hi = total if l == L - 1 else self._zt_offsets[l + 1][0][0][1]
                comm = self._comm_stream()
                comm.wait_event(main.record_event())
                for ev in events:
                    comm.wait_event(ev)
                with torch.cuda.stream(comm):
                    all_reduce.bucket(flat[lo:hi])

            def run_tail(l):
                """Parameter-only backward tails of layer l on their side streams (overlap the earlier layers' rows)."""
                n_before = len(tail_events)
                done = main.record_event()
                for c, side, ((Z, oZ), (th, oT)) in zip(self.conds[l], self._side_streams()[l], self._zt_offsets[l]):
                    if not self.params_on_side_streams:
                        side = main
                    else:
                        side.wait_event(done)
                    with torch.cuda.stream(side):
                        base = flat.data_ptr()
                        o_mu, o_sqrt, o_mA, o_mb = self._q_offsets[l]
                        self._timed("params_bwd", lambda: _cabi.check(lib.dgp_cond_backward_params(
                            c.ref(), _cabi.ptr(c.ws), klw, ctypes.c_void_p(base + 8 * oZ), ctypes.c_void_p(base + 8 * oT),
                            ctypes.c_void_p(base + 8 * o_mu), ctypes.c_void_p(base + 8 * o_sqrt),
                            ctypes.c_void_p(base + 8 * o_mA) if o_mA is not None else None,
                            ctypes.c_void_p(base + 8 * o_mb) if o_mb is not None else None, _cabi.stream_ptr()),
                            "dgp_cond_backward_params"))
                        if side is not main:
                            tail_events.append(side.record_event())
                if bucketed:
                    reduce_bucket(l, tail_events[n_before:])

            # offsets of each layer's gradient slots
            it = iter(zip(params, offs))
            self._zt_offsets, self._q_offsets = [], []
            for layer, conds in zip(self.layers, self.conds):
                self._zt_offsets.append([(next(it), next(it)) for _ in conds])
                (_, o_mu), (_, o_sqrt) = next(it), next(it)
                o_mA = o_mb = None
                if self.mean_trainable(layer):
                    (_, o_mA), (_, o_mb) = next(it), next(it)
                self._q_offsets.append((o_mu, o_sqrt, o_mA, o_mb))

            nchunks = (B + Bc - 1) // Bc
            for ck, b0 in enumerate(range(0, B, Bc)):
                b1 = min(B, b0 + Bc)
                nb = b1 - b0
                n = S * nb
                Xc, Yc = X[b0:b1], Y[b0:b1]
                buf = self._row_buffers(S, nb, need_grad)
                eps_c = None
                if eps_full is not None:
                    eps_c = [e[:, b0:b1].reshape(n, -1).contiguous() for e in eps_full]
                self._forward_rows(Xc, S, eps_c, seed, row_offset + b0, Bg, buf, need_grad, False, events)
                n_lik = self._rows_of(L - 1, S, nb)
                self._timed("lik", lambda: _cabi.check(lib.dgp_lik_gaussian(
                    _cabi.ptr(buf["mean"][-1]), _cabi.ptr(buf["var"][-1]), _cabi.ptr(Yc), nb, ldy, n_lik, RL,
                    _cabi.ptr(self._lik_var), T_lik, weight, ctypes.c_void_p(flat.data_ptr()),
                    _cabi.ptr(buf["g_mu"]) if need_grad else None, _cabi.ptr(buf["g_var"]) if need_grad else None,
                    ctypes.c_void_p(g_lik.data_ptr()) if need_grad else None,
                    _cabi.ptr(self._lik_ws), self._lik_ws.numel(), st), "dgp_lik_gaussian"))
                if ck == nchunks - 1:
                    # ELBO = weighted variational expectations - sum of the KL terms (every parameter stage has been waited
                    # for by the forward pass): final before the backward pass starts, so that the first layer's bucket
                    # can carry it
                    _cabi.check(lib.dgp_finish_elbo(ctypes.c_void_p(flat.data_ptr()), ctypes.c_void_p(kl.data_ptr()), ncond,
                                                    klw, st), "dgp_finish_elbo")
                if not need_grad:
                    continue
                for l in range(L - 1, -1, -1):
                    layer, conds = self.layers[l], self.conds[l]
                    inp = Xc if l == 0 else buf["F"][l - 1]
                    rows = self._rows_of(l, S, nb)
                    x_rows = nb if l <= 1 else n
                    if l == 0:
                        x_rows = nb
                    elif l == 1:
                        x_rows = n
                    last = (l == L - 1)
                    gF = None if last else buf["gX"][l + 1]
                    g_mean_in = buf["g_mu"] if last else None
                    g_var_in = buf["g_var"] if last else None
                    Fl = None if last else buf["F"][l]
                    if l == 0 and not last:
                        # fold the S sample gradients of every data row into (g_mean, g_var) of the shared conditional
                        self._timed("sample.l0", lambda: _cabi.check(lib.dgp_sample_reduce(
                            _cabi.ptr(gF), gF.shape[1], _cabi.ptr(buf["F"][0]), buf["F"][0].shape[1],
                            _cabi.ptr(buf["mean"][0]), _cabi.ptr(buf["var"][0]), S, nb, layer.output_dim,
                            layer.output_dim, _cabi.ptr(buf["g_mean0"]), _cabi.ptr(buf["g_var0"]), st),
                            "dgp_sample_reduce"))
                        g_mean_in, g_var_in, gF, Fl = buf["g_mean0"], buf["g_var0"], None, None
                    grouped = False
                    if l == 0 and self._groupable(conds):
                        # no input gradient at the data layer: the members' backward row kernels share one launch
                        n_c = len(conds)
                        descs = (ctypes.POINTER(_cabi.CondDesc) * n_c)(*[ctypes.pointer(c.desc) for c in conds])
                        wss = (ctypes.c_void_p * n_c)(*[c.ws.data_ptr() for c in conds])
                        sv = [(ctypes.c_void_p * n_c)(*[(b.data_ptr() if b is not None else None) for b in buf[k][l]])
                              for k in ("A", "U", "K")]
                        rc = self._timed(f"bwd.l{l}", lambda: lib.dgp_cond_backward_group(
                            n_c, descs, wss, _cabi.ptr(inp), x_rows, rows, sv[0], sv[1], sv[2], _cabi.ptr(buf["mean"][l]),
                            _cabi.ptr(buf["var"][l]), _cabi.ptr(Fl) if Fl is not None else None, buf["F"][l].shape[1],
                            _cabi.ptr(g_mean_in) if g_mean_in is not None else None,
                            _cabi.ptr(g_var_in) if g_var_in is not None else None,
                            _cabi.ptr(gF) if gF is not None else None, gF.shape[1] if gF is not None else 0,
                            1 if self.mean_trainable(layer) else 0, st))
                        if rc != _cabi.DGP_ERR_UNSUPPORTED:
                            _cabi.check(rc, "dgp_cond_backward_group")
                            grouped = True
                    for ci, c in enumerate(conds):
                        if not grouped:
                            self._timed(f"bwd.l{l}", lambda: _cabi.check(lib.dgp_cond_backward(
                                c.ref(), _cabi.ptr(c.ws), _cabi.ptr(inp), x_rows, rows, _cabi.ptr(buf["A"][l][ci]),
                                _cabi.ptr(buf["U"][l][ci]), _cabi.ptr(buf["K"][l][ci]), _cabi.ptr(buf["mean"][l]),
                                _cabi.ptr(buf["var"][l]),
                                _cabi.ptr(Fl) if Fl is not None else None, buf["F"][l].shape[1],
                                _cabi.ptr(g_mean_in) if g_mean_in is not None else None,
                                _cabi.ptr(g_var_in) if g_var_in is not None else None,
                                _cabi.ptr(gF) if gF is not None else None, gF.shape[1] if gF is not None else 0,
                                _cabi.ptr(buf["gX"][l]) if l > 0 else None, 1 if ci > 0 else 0,
                                1 if self.mean_trainable(layer) else 0, st), "dgp_cond_backward"))
                        if nchunks == 1 and l > 0 and self.gram_on_side_stream:
                            # the Gram reduction of this layer is not needed by the next layer's backward kernel: on the
                            # conditional's side stream it fills the SMs that kernel's last, partly empty wave leaves idle
                            side = self._side_streams()[l][ci]
                            side.wait_event(main.record_event())
                            with torch.cuda.stream(side):
                                self._timed(f"gram.l{l}", lambda: _cabi.check(lib.dgp_cond_backward_gram(
                                    c.ref(), _cabi.ptr(c.ws), _cabi.ptr(buf["A"][l][ci]), rows, 0, _cabi.stream_ptr()),
                                    "dgp_cond_backward_gram"))
                        else:
                            self._timed(f"gram.l{l}", lambda: _cabi.check(lib.dgp_cond_backward_gram(
                                c.ref(), _cabi.ptr(c.ws), _cabi.ptr(buf["A"][l][ci]), rows, 1 if ck > 0 else 0, st),
                                "dgp_cond_backward_gram"))
                    if ck == nchunks - 1:
                        run_tail(l)
            for ev in tail_events:
                main.wait_event(ev)
            if bucketed:
                all_reduce.finish()
            elif all_reduce is not None and world_size > 1:
                all_reduce(flat[:total])
            if check:
                self.check_status()
            grads = None
            if need_grad:
                grads = [flat[o:o + p.numel()].view(p.shape) for p, o in zip(params, offs)]
            self._last_flat = flat[:total]
            return flat[0], grads


class GraphedElbo:
    """One ELBO + gradient evaluation captured as a CUDA graph (fixed shapes): the ~100 launches of a step -- three
    streams' worth of parameter-stage kernels, the row-tile kernels, the tails -- are replayed with a single launch, which
    is what matters for small, launch-bound problems (BASELINE configs[0]) and trims the gaps of large ones.  Inputs are
    copied into static buffers before each replay; the Philox seed is added on the device (dgp_cond_forward's seed_dev)."""

    def __init__(self, engine, params, X, Y, S, num_data=None, row_offset=0, global_batch=None, world_size=1,
                 chunk_rows=None, need_grad=True, base_seed=0, all_reduce=None):
        """params: list of constrained parameter tensors copied into static buffers before every replay, or None to read
        the model's own parameters inside the graph (the positive transforms are then part of the capture; valid while
        the parameter tensors keep their storage, i.e. in-place optimiser updates)."""
        """all_reduce: a reducer (parallel.BucketReducer) whose NCCL collectives are captured INSIDE the graph, per-layer
        buckets overlapping the backward pass; None: the caller reduces the flat buffer after each replay (run())."""
        with torch.cuda.device(engine.device):
            self._build(engine, params, X, Y, S, num_data, row_offset, global_batch, world_size, chunk_rows, need_grad,
                        base_seed, all_reduce)

    @staticmethod
    def storage_key(engine):
        """Identity of the parameter storages a live-parameter graph reads (a re-assigned tensor needs a new capture)."""
        from .flatparams import FlatParams
        engine.flatten()
        return FlatParams.storage_key(engine)

    def _build(self, engine, params, X, Y, S, num_data, row_offset, global_batch, world_size, chunk_rows, need_grad,
               base_seed, all_reduce):
        self.eng = engine
        self.captured_collectives = all_reduce is not None
        self.live = params is None
        if self.live:
            with torch.no_grad():
                params = engine.params()
        dev = engine.device
        self.key = (tuple(X.shape), tuple(Y.shape), S, num_data, row_offset, global_batch, world_size, chunk_rows)
        self.Xs = _dev_tensor(X, dev).clone()
        self.Ys = _dev_tensor(Y, dev).clone()
        self.ps = None if self.live else [_dev_tensor(p, dev).clone() for p in params]
        live_ps = engine.native_params        # one transform launch (captured), then views of the model's own storage
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.base_seed = int(base_seed)
        kw = dict(S=S, eps=None, seed=base_seed, num_data=num_data, need_grad=need_grad, row_offset=row_offset,
                  global_batch=global_batch, world_size=world_size, all_reduce=all_reduce, chunk_rows=chunk_rows,
                  check=False)
        cur = torch.cuda.current_stream()
        if self.live:
            engine.flatten()                               # build the tables (allocations, H2D copies) outside the capture
        warm = torch.cuda.Stream(device=dev)
        warm.wait_stream(cur)
        with torch.cuda.stream(warm):                      # allocate every cached buffer before capture
            for _ in range(2):
                with torch.no_grad():
                    engine.elbo(live_ps() if self.live else self.ps, self.Xs, self.Ys, **kw)
        cur.wait_stream(warm)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        engine._seed_dev = self.seed_dev
        try:
            # thread-local capture mode: NCCL's watchdog thread may touch the CUDA API while this thread captures
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                with torch.no_grad():
                    self.val, self.grads = engine.elbo(live_ps() if self.live else self.ps, self.Xs, self.Ys, **kw)
                self.flat = engine._last_flat
        finally:
            engine._seed_dev = None
        # the capture holds raw device pointers into the engine's workspaces and row buffers: keep them alive for as long
        # as this graph lives, whatever the engine allocates (or drops from its caches) for later calls
        self._pins = ([c.ws for conds in engine.conds for c in conds], dict(engine._rows), engine._lik_ws, engine._flat)

    def run(self, params, X, Y, seed, all_reduce=None):
        """Replay with the given parameters, minibatch and (full) Philox seed; returns views of static buffers."""
        with torch.no_grad(), torch.cuda.device(self.eng.device):
            if not self.live:
                torch._foreach_copy_(self.ps, [p.detach() for p in params])
            self.Xs.copy_(X, non_blocking=True)
            self.Ys.copy_(Y, non_blocking=True)
            # Philox seed offset (added to base_seed on the device): a stream-ordered device write in front of the replay,
            # so that back-to-back replays each see their own seed
            _cabi.check(_cabi.lib().dgp_set_u64(ctypes.c_void_p(self.seed_dev.data_ptr()),
                                                (int(seed) - self.base_seed) & 0x7FFFFFFFFFFFFFFF, _cabi.stream_ptr()),
                        "dgp_set_u64")
            self.graph.replay()
            if all_reduce is not None and not self.captured_collectives:
                all_reduce(self.flat)
        return self.val, self.grads


class _ElboFn(torch.autograd.Function):
    """ELBO as one autograd node: forward runs the fused forward AND backward kernels; backward scales the stored
    gradients.  The softplus chain of positive parameters is left to torch (it sits outside this node)."""

    @staticmethod
    def forward(ctx, engine, kwargs, *params):
        need = any(ctx.needs_input_grad[2:])
        kwargs = dict(kwargs)
        graph = kwargs.pop("_graph", None)
        if graph is not None and need:
            val, grads = graph.run(list(params), kwargs["X"], kwargs["Y"], kwargs["seed"],
                                   all_reduce=kwargs.get("all_reduce"))
            if kwargs.get("check"):
                engine.check_status()
            grads = [g.clone() for g in grads]      # the graph's output buffers are overwritten by the next replay
        else:
            val, grads = engine.elbo(list(params), need_grad=need, **kwargs)
        ctx.grads = grads
        return val.clone()

    @staticmethod
    def backward(ctx, g):
        if ctx.grads is None:
            return (None, None) + tuple(None for _ in ctx.needs_input_grad[2:])
        return (None, None) + tuple(g * gr for gr in ctx.grads)


def elbo_autograd(engine, **kwargs):
    return _ElboFn.apply(engine, kwargs, *engine.params())


# ---------------------------------------------------------------------------------------------- single-layer helpers
def _layer_engine(layer):
    eng = getattr(layer, "_dgp_engine", None)
    if eng is None or eng.device.index != torch.cuda.current_device():
        eng = Engine([layer], None, multitask=False)
        object.__setattr__(layer, "_dgp_engine", eng)
    return eng


def layer_predict_full_cov(layer, Xnew, stochastic=True):
    """Layer._build_predict(full_cov=True) (dgplib/layers.py:74-81): stochastic: [S,N,D] -> mean [S,N,R], var [S,N,N,R];
    otherwise [N,D] -> mean [N,R], var [R,N,N] (gpflow's own layout, as the reference returns it there)."""
    eng = _layer_engine(layer)
    lib, st = _cabi.lib(), _cabi.stream_ptr()
    with torch.no_grad():
        X = _dev_tensor(Xnew, eng.device)
        eng._bind(eng._detached())
        rows = X.shape[1] if stochastic else X.shape[0]
        for evs in eng._prepare(1, rows, False, None):
            for ev in evs:
                torch.cuda.current_stream().wait_event(ev)
        if not stochastic:
            return eng._layer_full_cov(0, X, lib, st)
        means, vars_ = [], []
        for s in range(X.shape[0]):
            m, v = eng._layer_full_cov(0, X[s].contiguous(), lib, st)
            means.append(m)
            vars_.append(v.permute(2, 1, 0).contiguous())
        return torch.stack(means), torch.stack(vars_)


def layer_predict(layer, Xnew, stochastic=True):
    """Layer._build_predict (dgplib/layers.py:62-94), diag branch.  Returns CUDA tensors."""
    eng = _layer_engine(layer)
    X = _dev_tensor(Xnew, eng.device)
    if stochastic:
        S, N, D = X.shape
        flat = X.reshape(S * N, D)          # dgplib/layers.py:84-85
    else:
        S, (N, D) = 1, X.shape
        flat = X
    _, means, vars_ = eng.propagate(flat, 1, eps=[torch.zeros(S * N, layer.output_dim, dtype=_F64)])
    mean, var = means[0][0], vars_[0][0]
    if stochastic:
        return mean.reshape(S, N, -1), var.reshape(S, N, -1)
    return mean, var


def layer_kl(layer):
    eng = _layer_engine(layer)
    with torch.no_grad():
        eng._bind(eng._detached())
        ncond = len(eng.conds[0])
        kl = torch.zeros(ncond, dtype=_F64, device=eng.device)
        for evs in eng._prepare(1, 1, False, kl):
            for ev in evs:
                torch.cuda.current_stream().wait_event(ev)
        return kl.sum()


def _kernel_desc(kernel, Dx, device):
    parts = kernel.task_kernels()
    d = _cabi.CondDesc()
    d.T = len(parts)
    d.Dx = Dx
    d.D = min(kernel.input_dim, Dx - 1) if kernel.switched else kernel.input_dim
    d.M = d.R = d.ldr = 1
    for t, p in enumerate(parts):
        d.kind[t] = p[0]
    rows = []
    for (_, var, white, ls) in parts:
        w = torch.zeros((), dtype=_F64) if white is None else white
        rows.append(torch.cat([var.reshape(1).cpu(), w.reshape(1).cpu(), ls[:d.D].cpu()]))
    theta = _dev_tensor(torch.stack(rows), device)
    d.theta = theta.data_ptr()
    return d, theta


def kernel_K(kernel, X, X2=None):
    """gpflow Kernel.K(X, X2) on the GPU (dgp_kernel_K)."""
    _cabi.require_cuda()
    dev = torch.device(f"cuda:{torch.cuda.current_device()}")
    X = _dev_tensor(X, dev)
    d, theta = _kernel_desc(kernel, X.shape[1], dev)
    X2d = None if X2 is None else _dev_tensor(X2, dev)
    n2 = X.shape[0] if X2 is None else X2d.shape[0]
    K = torch.empty(X.shape[0], n2, dtype=_F64, device=dev)
    _cabi.check(_cabi.lib().dgp_kernel_K(ctypes.byref(d), _cabi.ptr(X), X.shape[0], _cabi.ptr(X2d), n2,
                                         1 if X2 is None else 0, _cabi.ptr(K), _cabi.stream_ptr()), "dgp_kernel_K")
    return K


def kernel_Kdiag(kernel, X):
    _cabi.require_cuda()
    dev = torch.device(f"cuda:{torch.cuda.current_device()}")
    X = _dev_tensor(X, dev)
    d, theta = _kernel_desc(kernel, X.shape[1], dev)
    out = torch.empty(X.shape[0], dtype=_F64, device=dev)
    _cabi.check(_cabi.lib().dgp_kernel_Kdiag(ctypes.byref(d), _cabi.ptr(X), X.shape[0], _cabi.ptr(out),
                                             _cabi.stream_ptr()), "dgp_kernel_Kdiag")
    return out


def pca_weights(X, output_dim, return_singular_values=False):
    """Leading `output_dim` right-singular vectors of X [N, D] as a NumPy [D, output_dim] matrix, computed on the GPU
    (dgp_pca_weights, csrc/pca.cu) -- the SVD step of find_weights (dgplib/layers.py:104-109)."""
    _cabi.require_cuda()
    lib = _cabi.lib()
    dev = torch.device("cuda", torch.cuda.current_device())
    Xd = X if isinstance(X, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64))
    Xd = Xd.to(device=dev, dtype=torch.float64).contiguous()
    n, D = Xd.shape
    wsb = lib.dgp_pca_ws_bytes(n, D)
    if wsb == 0:
        raise _cabi.DgpError(f"dgp_pca_ws_bytes: unsupported shape N={n}, D={D} (D <= {_cabi.DGP_MAX_D})")
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    W = torch.empty(D, output_dim, dtype=torch.float64, device=dev)
    sing = torch.empty(D, dtype=torch.float64, device=dev)
    _cabi.check(lib.dgp_pca_weights(_cabi.ptr(Xd), n, D, D, output_dim, _cabi.ptr(ws), wsb, _cabi.ptr(W), _cabi.ptr(sing), _cabi.stream_ptr()),
            "dgp_pca_weights")
    out = W.cpu().numpy()
    return (out, sing.cpu().numpy()) if return_singular_values else out
