"""Layer classes of the DGP cascade -- same surface as dgplib/layers.py.

`Layer` (alias `SVGP_Layer`) owns Z (in `feature`), q_mu, q_sqrt, the kernel and the mean function
(dgplib/layers.py:21-56); `_build_predict` evaluates the whitened sparse-GP conditional with the fused sm_100a kernels
(csrc/layer_fwd.cu) instead of gpflow.conditionals.conditional (dgplib/layers.py:62-94); `build_prior_KL` is the
whitened gauss_kl (dgplib/layers.py:58-60).  The Input/Hidden/Output mixins and `initialize_forward` are host-side
initialisation (dgplib/layers.py:97-218), NumPy in / NumPy out, exactly as in the reference.
"""
import numpy as np

from . import settings
from .features import inducingpoint_wrapper
from .mean_functions import Linear, Zero
from .params import Param, ParamList, Parameterized


class Layer(Parameterized):
    """The basic layer class.  Handles input_dim and output_dim."""

    def __init__(self, input_dim, output_dim, num_inducing, kernel, mean_function=None, multitask=False, name=None):
        super().__init__(name=name)
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.num_inducing = num_inducing
        self.multitask = bool(multitask)
        Z = np.zeros((self.num_inducing, self.input_dim + (1 if multitask else 0)))   # dgplib/layers.py:36-39
        self.feature = inducingpoint_wrapper(None, Z)
        self.kernel = ParamList(kernel) if isinstance(kernel, list) else kernel      # :43-46
        self.mean_function = mean_function or Zero(output_dim=self.output_dim)       # :48
        self.q_mu = Param(np.zeros((self.num_inducing, self.output_dim)))            # :50-52
        self.q_sqrt = Param(np.tile(np.eye(self.num_inducing)[None], (self.output_dim, 1, 1)))  # :54-56

    # conditionals of this layer: list of (kernel, feature, first output column, number of outputs)
    def conditionals(self):
        return [(self.kernel, self.feature, 0, self.output_dim)]

    def build_prior_KL(self, K=None):
        """gauss_kl(q_mu, q_sqrt, K=None) (dgplib/layers.py:58-60); returns a CUDA scalar tensor."""
        if K:
            raise NotImplementedError("only the whitened KL (K=None) is on the DSDGP path (dgplib/dsdgp.py:126)")
        from .engine import layer_kl
        return layer_kl(self)

    def _build_predict(self, Xnew, full_cov=False, stochastic=True):
        """(mean, var) of the layer at Xnew.  stochastic=True: Xnew [S,N,D] -> [S,N,R] each (samples flattened into
        one conditional, dgplib/layers.py:82-87); stochastic=False: [N,D] -> [N,R].  full_cov=True (N <= 16384):
        var is [S,N,N,R] (dgplib/layers.py:74-81), or gpflow's [R,N,N] when stochastic=False."""
        if full_cov:
            from .engine import layer_predict_full_cov
            return layer_predict_full_cov(self, Xnew, stochastic)
        from .engine import layer_predict
        return layer_predict(self, Xnew, stochastic)


SVGP_Layer = Layer   # the name BASELINE.json's north_star uses for this class


def find_weights(input_dim, output_dim, X, multitask=False, backend=None):
    """Initial weights of a layer's Linear mean function (dgplib/layers.py:97-121): identity when the widths agree,
    the leading right-singular vectors of X when the layer narrows, identity padded with zero columns when it widens;
    multitask appends a zero row so that the task-id column does not leak into the mean.
    backend (default settings.pca_backend): "numpy" = host LAPACK SVD as in the reference, "cuda" = dgp_pca_weights."""
    if input_dim == output_dim:
        W = np.identity(input_dim)
    elif input_dim > output_dim:
        data = X[:, :-1] if multitask else X
        backend = backend or settings.pca_backend
        if backend == "cuda":
            from .engine import pca_weights
            W = pca_weights(data, output_dim)
        elif backend == "numpy":
            _, _, Vt = np.linalg.svd(data, full_matrices=False)
            W = Vt[:output_dim].T
        else:
            raise ValueError(f"unknown pca backend {backend!r} (numpy | cuda)")
    else:
        W = np.zeros((input_dim, output_dim))
        W[:, :input_dim] = np.identity(input_dim)
    if multitask:
        W = np.concatenate([W, np.zeros((1, W.shape[1]))], axis=0)
    return W


def _first_Z(feature):
    feat = feature[0] if isinstance(feature, ParamList) else feature
    return feat.Z.value


class InputMixin(object):
    """Computes the inputs / inducing inputs handed to the next layer (dgplib/layers.py:124-136)."""

    def compute_inputs(self, X, Z, multitask=False):
        W = find_weights(self.input_dim, self.output_dim, X, multitask)
        return X.copy().dot(W), Z.copy().dot(W), W


class HiddenMixin(object):
    def compute_inputs(self, X, Z, multitask=False):                 # dgplib/layers.py:139-154
        W = find_weights(self.input_dim, self.output_dim, X, multitask)
        return X.copy().dot(W), _first_Z(self.feature).dot(W), W


class OutputMixin(object):
    def compute_inputs(self, X, Z, multitask=False):                 # dgplib/layers.py:157-170
        W = find_weights(self.input_dim, self.output_dim, X, multitask)
        return X.copy().dot(W), _first_Z(self.feature).dot(W), W


def _assign_Z(layer, Z):
    feats = layer.feature if isinstance(layer.feature, ParamList) else [layer.feature]
    for feat in feats:
        feat.Z.assign(Z)


def _init_forward(layer, X, Z, multitask):
    """Shared body of Input/Hidden initialize_forward (dgplib/layers.py:175-208)."""
    _assign_Z(layer, Z)
    X_running, Z_running, W = layer.compute_inputs(X, Z, multitask)
    if isinstance(layer.mean_function, Linear):
        layer.mean_function.A = W
        layer.mean_function.set_trainable(False)
    return X_running, Z_running


class InputLayer(Layer, InputMixin):
    def initialize_forward(self, X, Z, multitask=False):
        return _init_forward(self, X, Z, multitask)


class HiddenLayer(Layer, HiddenMixin):
    def initialize_forward(self, X, Z, multitask=False):
        return _init_forward(self, X, Z, multitask)


class OutputLayer(Layer, OutputMixin):
    def initialize_forward(self, X, Z, multitask=False):              # dgplib/layers.py:211-218
        _assign_Z(self, Z)
        return (None, None)
