"""gpflow.kernels stand-ins on the DSDGP path: RBF, Matern52, White and their Sum.  These objects
only hold parameters; the covariance arithmetic is CUDA (csrc/layer_fwd.cu phase K, csrc/prep.cu k_kuu,
csrc/likelihood.cu k_kernel_K).  `active_dims` is always slice(input_dim), gpflow's default, which is what makes a
plain kernel ignore a trailing task-id column."""
import numpy as np
import torch

from . import _cabi
from .params import Param, Parameterized, ParamList


class Kernel(Parameterized):
    def __init__(self, input_dim, name=None):
        super().__init__(name)
        self.input_dim = int(input_dim)

    def __add__(self, other):
        return Sum([self, other])

    # -- what the CUDA path needs: (kind, variance[], white[] or None, lengthscales[D]) per task kernel
    def stationary_parts(self):
        raise NotImplementedError(f"{type(self).__name__} is not supported by the B200 DSDGP path "
                                  "(supported: RBF or Matern52, optionally + White, optionally inside a SwitchedKernel)")

    def task_kernels(self):
        """List of per-task (kind, variance, white, lengthscales) tuples; one entry for a non-switched kernel."""
        return [self.stationary_parts()]

    # -- the Param objects behind stationary_parts(): (variance Param, [White variance Params], lengthscales Param);
    #    flatparams.py builds the transform / chain-rule tables of the native training step from them
    def param_refs(self):
        raise NotImplementedError(f"{type(self).__name__} is not supported by the B200 DSDGP path")

    def task_param_refs(self):
        return [self.param_refs()]

    @property
    def switched(self):
        return False

    # -- stand-alone gram matrices (gpflow Kernel.K / Kdiag) through the C-ABI
    def K(self, X, X2=None):
        from .engine import kernel_K
        return kernel_K(self, X, X2)

    def Kdiag(self, X):
        from .engine import kernel_Kdiag
        return kernel_Kdiag(self, X)


class Stationary(Kernel):
    KIND = None

    def __init__(self, input_dim, variance=1.0, lengthscales=1.0, ARD=False, name=None):
        super().__init__(input_dim, name)
        self.ARD = bool(ARD) or np.ndim(lengthscales) > 0
        if self.ARD:
            ls = np.ones(self.input_dim) * np.asarray(lengthscales, dtype=np.float64)
        else:
            ls = float(lengthscales)
        self.variance = Param(variance, positive=True)
        self.lengthscales = Param(ls, positive=True)

    def stationary_parts(self):
        ls = self.lengthscales.constrained()
        ls = ls.expand(self.input_dim) if ls.dim() == 0 else ls
        return (self.KIND, self.variance.constrained(), None, ls)

    def param_refs(self):
        return (self.variance, [], self.lengthscales)


class RBF(Stationary):
    KIND = _cabi.DGP_KIND_RBF


class Matern52(Stationary):
    KIND = _cabi.DGP_KIND_MATERN52


class White(Kernel):
    def __init__(self, input_dim, variance=1.0, name=None):
        super().__init__(input_dim, name)
        self.variance = Param(variance, positive=True)


class Combination(Kernel):
    def __init__(self, kernels, name=None):
        input_dim = max(k.input_dim for k in kernels)
        super().__init__(input_dim, name)
        flat = []
        for k in kernels:
            if type(self) is Sum and isinstance(k, Sum):
                flat.extend(k.kernels)
            else:
                flat.append(k)
        self.kernels = ParamList(flat)


class Sum(Combination):
    def stationary_parts(self):
        stat = [k for k in self.kernels if isinstance(k, Stationary)]
        white = [k for k in self.kernels if isinstance(k, White)]
        if len(stat) != 1 or len(stat) + len(white) != len(self.kernels):
            raise NotImplementedError("the B200 DSDGP path supports sums of exactly one RBF/Matern52 and White kernels")
        kind, var, _, ls = stat[0].stationary_parts()
        w = None
        for k in white:
            w = k.variance.constrained() if w is None else w + k.variance.constrained()
        return (kind, var, w, ls)

    def param_refs(self):
        self.stationary_parts()      # raises for unsupported sums
        stat = [k for k in self.kernels if isinstance(k, Stationary)][0]
        return (stat.variance, [k.variance for k in self.kernels if isinstance(k, White)], stat.lengthscales)

    @property
    def input_dim_active(self):
        return self.input_dim
