"""gpflow.mean_functions stand-ins used by the reference: Zero (dgplib/layers.py:48) and Linear (dgplib/layers.py:184-186,
whose A is overwritten by find_weights and frozen in initialize_forward)."""
import numpy as np

from .params import Param, Parameterized


class MeanFunction(Parameterized):
    pass


class Zero(MeanFunction):
    def __init__(self, output_dim=1, name=None):
        super().__init__(name)
        self.output_dim = output_dim


class Linear(MeanFunction):
    """X @ A + b."""

    def __init__(self, A=None, b=None, name=None):
        super().__init__(name)
        A = np.ones((1, 1)) if A is None else np.atleast_2d(np.asarray(A, dtype=np.float64))
        b = np.zeros(A.shape[1]) if b is None else np.atleast_1d(np.asarray(b, dtype=np.float64))
        self._A = Param(A)
        self._b = Param(b)

    @property
    def A(self):
        return self._A

    @A.setter
    def A(self, value):          # `mean_function.A = W` in the reference (dgplib/layers.py:185)
        if isinstance(value, Param):
            torch_module_setattr(self, "_A", value)
        else:
            self._A.assign(value)

    @property
    def b(self):
        return self._b

    @b.setter
    def b(self, value):
        if isinstance(value, Param):
            torch_module_setattr(self, "_b", value)
        else:
            self._b.assign(value)


def torch_module_setattr(mod, name, value):
    import torch
    torch.nn.Module.__setattr__(mod, name, value)
