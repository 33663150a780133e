"""Parameter plumbing: light stand-ins for gpflow.params.{Parameter, Parameterized, ParamList} built on
torch.nn.  Positive parameters use gpflow's Log1pe transform, theta = softplus(u) + 1e-6, and optimisers see the
unconstrained u (the reference differentiates w.r.t. unconstrained variables)."""
import numpy as np
import torch

from . import settings


# Bumped whenever the SET of torch Parameters of any model can have changed (a Param created or its `raw` replaced, a
# sub-module attached or replaced, a ParamList edited).  The engine caches its walk over the module tree on it; storage
# addresses and trainable flags are still compared on every evaluation (flatparams.FlatParams.storage_key).
_structure_epoch = [0]


def structure_epoch():
    return _structure_epoch[0]


def _structure_changed():
    _structure_epoch[0] += 1


def _to_tensor(value):
    if isinstance(value, torch.Tensor):
        return value.detach().to(settings.float_type).clone()
    return torch.tensor(np.array(value, dtype=np.float64), dtype=settings.float_type)


class Param(torch.nn.Module):
    """One (optionally positive-constrained) float64 parameter."""

    def __init__(self, value, positive=False, trainable=True):
        super().__init__()
        self.positive = bool(positive)
        self.raw = torch.nn.Parameter(self._backward(_to_tensor(value)), requires_grad=bool(trainable))

    def __setattr__(self, name, value):
        if isinstance(value, (torch.nn.Parameter, torch.nn.Module)):
            _structure_changed()
        super().__setattr__(name, value)

    def _backward(self, y):
        if not self.positive:
            return y
        s = y - settings.positive_lower
        if bool((s <= 0).any()):
            raise ValueError("positive parameter must exceed %g" % settings.positive_lower)
        return s + torch.log(-torch.expm1(-s))      # inverse softplus, stable

    def constrained(self):
        """Constrained value as a tensor attached to the autograd graph."""
        if not self.positive:
            return self.raw
        return torch.nn.functional.softplus(self.raw) + settings.positive_lower

    @property
    def value(self):
        return self.constrained().detach().cpu().numpy().copy()

    def read_value(self):
        return self.value

    @property
    def shape(self):
        return tuple(self.raw.shape)

    @property
    def trainable(self):
        return self.raw.requires_grad

    def set_trainable(self, flag):
        self.raw.requires_grad_(bool(flag))

    def assign(self, value):
        v = _to_tensor(value).to(self.raw.device)
        if tuple(v.shape) != tuple(self.raw.shape):
            # gpflow allows shape changes before the graph is built (Linear.A is replaced in initialize_forward)
            self.raw = torch.nn.Parameter(self._backward(v), requires_grad=self.raw.requires_grad)
        else:
            with torch.no_grad():
                self.raw.copy_(self._backward(v))


class Parameterized(torch.nn.Module):
    """Base of kernels, likelihoods, mean functions, layers and models (gpflow.params.Parameterized)."""

    def __init__(self, name=None):
        super().__init__()
        self.name = name or type(self).__name__

    def __setattr__(self, name, value):
        if isinstance(value, (torch.nn.Parameter, torch.nn.Module)):
            _structure_changed()
        super().__setattr__(name, value)

    def __delattr__(self, name):
        _structure_changed()
        super().__delattr__(name)

    def set_trainable(self, flag):
        for p in self.parameters():
            p.requires_grad_(bool(flag))

    def trainable_parameters(self):
        return [p for p in self.parameters() if p.requires_grad]


class ParamList(torch.nn.ModuleList):
    """gpflow.params.ParamList: an indexable list of Parameterized objects."""

    def __bool__(self):
        return len(self) > 0

    def __setitem__(self, idx, module):
        _structure_changed()
        return super().__setitem__(idx, module)

    def __delitem__(self, idx):
        _structure_changed()
        return super().__delitem__(idx)

    def append(self, module):
        _structure_changed()
        return super().append(module)

    def extend(self, modules):
        _structure_changed()
        return super().extend(modules)

    def insert(self, index, module):
        _structure_changed()
        return super().insert(index, module)
