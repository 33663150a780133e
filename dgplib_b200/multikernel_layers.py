"""MultikernelLayer -- same surface as dgplib/multikernel_layers.py: one whitened conditional per kernel on a column
slice of q_mu / q_sqrt (offset = output_dim / num_kernels), Z shared or one copy per kernel."""
import numpy as np

from .features import inducingpoint_wrapper
from .layers import Layer, InputMixin, HiddenMixin, OutputMixin, _init_forward, _assign_Z
from .params import ParamList


class MultikernelLayer(Layer):
    """Inherits from Layer.  Can handle outputs from different priors."""

    def __init__(self, input_dim, output_dim, num_inducing, kernel_list, share_Z=False, mean_function=None,
                 multitask=False, name=None):
        if output_dim % len(kernel_list) != 0:                       # dgplib/multikernel_layers.py:27-28
            raise ValueError("Output dimension must be a multiple of the number of kernels")
        super().__init__(input_dim=input_dim, output_dim=output_dim, num_inducing=num_inducing,
                         kernel=list(kernel_list), mean_function=mean_function, multitask=multitask, name=name)
        self.num_kernels = len(kernel_list)
        self._shared_Z = share_Z
        self.offset = int(self.output_dim / self.num_kernels)        # :40
        if not self._shared_Z:                                        # :42-49
            Z = np.zeros((self.num_inducing, self.input_dim + (1 if multitask else 0)))
            del self.feature
            self.feature = ParamList([inducingpoint_wrapper(None, Z.copy()) for _ in range(self.num_kernels)])

    def conditionals(self):
        feats = [self.feature] * self.num_kernels if self._shared_Z else list(self.feature)   # :69-73
        return [(k, f, i * self.offset, self.offset) for i, (k, f) in enumerate(zip(self.kernel, feats))]

    def build_prior_KL(self, K=None):
        # the reference's `if K:` branch calls an undefined gauss_kl_white (dgplib/multikernel_layers.py:54-61) and is
        # never reached (dgplib/dsdgp.py:126 passes K=None); only the whitened KL exists.
        return super().build_prior_KL(K)


class MultikernelInputLayer(MultikernelLayer, InputMixin):
    def initialize_forward(self, X, Z, multitask=False):
        return _init_forward(self, X, Z, multitask)


class MultikernelHiddenLayer(MultikernelLayer, HiddenMixin):
    def initialize_forward(self, X, Z, multitask=False):
        return _init_forward(self, X, Z, multitask)


class MultikernelOutputLayer(MultikernelLayer, OutputMixin):
    def initialize_forward(self, X, Z, multitask=False):
        _assign_Z(self, Z)
        return (None, None)
