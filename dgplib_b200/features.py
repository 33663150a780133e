"""gpflow.features.InducingPoints stand-in: the object the reference keeps in Layer.feature (dgplib/layers.py:41)."""
from .params import Param, Parameterized


class InducingPoints(Parameterized):
    def __init__(self, Z, name=None):
        super().__init__(name)
        self.Z = Param(Z)

    def __len__(self):
        return self.Z.shape[0]


def inducingpoint_wrapper(feat, Z):
    """gpflow.features.inducingpoint_wrapper(None, Z) as called at dgplib/layers.py:41."""
    if feat is not None and Z is not None:
        raise ValueError("Cannot pass both an InducingFeature instance and Z values")
    return feat if feat is not None else InducingPoints(Z)
