"""Sequential / MultitaskSequential -- same surface and validation rules as dgplib/cascade.py (host-side container;
no arithmetic on the per-step path)."""
import numpy as np

from .layers import InputMixin, OutputMixin, Layer
from .layers import InputLayer, OutputLayer, HiddenLayer  # noqa: F401  (re-exported like the reference)
from .multikernel_layers import (MultikernelLayer, MultikernelInputLayer, MultikernelOutputLayer,  # noqa: F401
                                 MultikernelHiddenLayer)
from .params import ParamList, Parameterized


class Sequential(Parameterized):
    """Linear stack of layers."""

    def __init__(self, layers=None, name=None):
        super().__init__(name=name)
        self._initialized = False
        self.layers = ParamList([])
        for layer in layers or []:
            self.add(layer)

    def _valid_input(self, layer):
        assert isinstance(layer, Layer)
        if self._initialized:
            raise ValueError('Cannot add more layers to initialized model')
        if not self.layers:
            assert isinstance(layer, InputMixin), "First layer must be an Input Layer"
        else:
            if isinstance(self.layers[-1], OutputMixin):
                raise ValueError('Cannot add layers after an Output Layer')
            assert self.layers[-1].output_dim == layer.input_dim, \
                "Input dimensions of layer must be equal to the output dimensions of the preceding layer"

    def add(self, layer):
        """Adds a layer instance on top of the layer stack."""
        self._valid_input(layer)
        self.layers.append(layer)

    def get_dims(self):
        return [(l.input_dim, l.output_dim) for l in self.layers]

    def initialize_params(self, X, Z):
        X_running, Z_running = self.layers[0].initialize_forward(X, Z)
        for layer in list(self.layers)[1:]:
            X_running, Z_running = layer.initialize_forward(X_running, Z_running)
        print('Model Parameters Initialized')
        self._initialized = True


class MultitaskSequential(Sequential):
    def initialize_params(self, X, Z):
        X_ind, Z_ind = X[:, -1:], Z[:, -1:]                           # dgplib/cascade.py:72-74
        X_running, Z_running = self.layers[0].initialize_forward(X, Z, multitask=True)
        for layer in list(self.layers)[1:]:
            X_running = np.hstack([X_running, X_ind])
            Z_running = np.hstack([Z_running, Z_ind])
            X_running, Z_running = layer.initialize_forward(X_running, Z_running, multitask=True)
        print('Model Parameters Initialized')
        self._initialized = True
