"""Layer containers with the surface of dgplib/cascade.py: `Sequential(layers=None)`, `.add(layer)`, `.get_dims()`,
`.initialize_params(X, Z)`, `._initialized`, and `MultitaskSequential` (the task-id column rides beside every layer's
inputs).  Host-side bookkeeping only; nothing here is on the per-step path.

Validation contract (dgplib/cascade.py:28-46, pinned by the reference's tests/test_cascade.py:33-37, :103-122):
AssertionError for something that is not a Layer, for a first layer that is not an input layer and for a width mismatch;
ValueError for adding to an initialised stack or on top of an output layer."""
import numpy as np

from .layers import InputMixin, OutputMixin, Layer
from .layers import InputLayer, OutputLayer, HiddenLayer  # noqa: F401  (importable from here, as in the reference)
from .multikernel_layers import (MultikernelLayer, MultikernelInputLayer, MultikernelOutputLayer,  # noqa: F401
                                 MultikernelHiddenLayer)
from .params import ParamList, Parameterized


class Sequential(Parameterized):
    """An ordered stack of DGP layers: input layer first, optionally hidden layers, an output layer last."""

    multitask = False      # MultitaskSequential: carry the last column (task id) around every layer

    def __init__(self, layers=None, name=None):
        super().__init__(name=name)
        self._initialized = False
        self.layers = ParamList([])
        for layer in (layers or ()):
            self.add(layer)

    # -- construction -------------------------------------------------------------------------------------------
    def _why_not(self, layer):
        """(exception type, message) if `layer` may not be appended now, else None."""
        if not isinstance(layer, Layer):
            return AssertionError, f"expected a Layer, got {type(layer).__name__}"
        if self._initialized:
            return ValueError, "the stack has been initialised; layers can no longer be added"
        if len(self.layers) == 0:
            if not isinstance(layer, InputMixin):
                return AssertionError, "the first layer of a stack must be an input layer"
            return None
        top = self.layers[-1]
        if isinstance(top, OutputMixin):
            return ValueError, "the stack already ends in an output layer"
        if top.output_dim != layer.input_dim:
            return AssertionError, (f"width mismatch: the stack ends with {top.output_dim} outputs, "
                                    f"the new layer takes {layer.input_dim} inputs")
        return None

    def add(self, layer):
        """Append `layer` to the stack (validated against the layers already there)."""
        problem = self._why_not(layer)
        if problem is not None:
            raise problem[0](problem[1])
        self.layers.append(layer)

    def get_dims(self):
        """[(input_dim, output_dim)] of the layers, bottom to top."""
        return [(layer.input_dim, layer.output_dim) for layer in self.layers]

    # -- initialisation (dgplib/cascade.py:61-67, :70-82) ---------------------------------------------------------
    def initialize_params(self, X, Z):
        """Push the data X and the inducing inputs Z through every layer's `initialize_forward`, which sets the layer's
        inducing points and its Linear mean weights and returns the inputs of the next layer."""
        ids = (X[:, -1:], Z[:, -1:]) if self.multitask else None
        extra = {"multitask": True} if self.multitask else {}
        running = (X, Z)
        for depth, layer in enumerate(self.layers):
            if depth > 0 and ids is not None:      # re-attach the task-id column dropped by the previous layer's output
                running = (np.hstack([running[0], ids[0]]), np.hstack([running[1], ids[1]]))
            running = layer.initialize_forward(running[0], running[1], **extra)
        self._initialized = True
        print("Model Parameters Initialized")


class MultitaskSequential(Sequential):
    """Sequential for inputs whose last column is an integer task id (dgplib/cascade.py:69-82)."""

    multitask = True
