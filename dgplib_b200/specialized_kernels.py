"""dgplib/specialized_kernels.py surface: `SwitchedKernel`, plus the entry points of the hand-written sm_100a
covariance kernels that BASELINE.json's north_star places in this module (`kuf`, `kuu`, `kdiag`).

SwitchedKernel picks the kernel by the integer task id in the last input column: K[i, j] = k_t(x_i, x2_j) when
task_i == task_j == t and 0 otherwise (dgplib/specialized_kernels.py:23-60); Kdiag[i] = k_{task_i}.Kdiag
(:62-72).  On the GPU this is a per-row hyper-parameter select plus an equality mask inside the Kuf / Kuu kernels --
no partition / gather / scatter / stitch."""
from .kernels import Combination, Kernel


class SwitchedKernel(Combination):
    def __init__(self, kern_list, output_dim, name=None):
        super().__init__(kernels=list(kern_list), name=name)
        self.output_dim = output_dim
        self.num_kernels = len(self.kernels)
        assert self.output_dim == self.num_kernels                   # dgplib/specialized_kernels.py:21

    @property
    def switched(self):
        return True

    def task_kernels(self):
        return [k.stationary_parts() for k in self.kernels]

    def task_param_refs(self):
        return [k.param_refs() for k in self.kernels]


def kuf(kernel, Z, X):
    """Kuf = k(Z, X) [M, N] (gpflow features.Kuf): CUDA, through dgp_kernel_K."""
    return kernel.K(Z, X)


def kuu(kernel, Z, jitter=0.0):
    """Kuu = k(Z, Z) + jitter I [M, M] (gpflow features.Kuu)."""
    import torch
    K = kernel.K(Z)
    return K + jitter * torch.eye(K.shape[0], dtype=K.dtype, device=K.device)


def kdiag(kernel, X):
    """diag k(X, X) [N] (gpflow Kernel.Kdiag)."""
    return kernel.Kdiag(X)


__all__ = ["SwitchedKernel", "kuf", "kuu", "kdiag", "Kernel"]
