"""dgplib_b200 -- B200-native (sm_100a) implementation of dgplib's doubly stochastic deep GP path.

Mirrors the reference package layout (dgplib/__init__.py): submodules `layers`, `cascade`, `multikernel_layers`,
`utils`, `specialized_kernels`, and the models `DSDGP`, `MultitaskDSDGP`; plus light stand-ins for the GPflow objects
the reference's users construct (`kernels`, `likelihoods`, `mean_functions`).
"""
from . import settings
from . import kernels
from . import likelihoods
from . import mean_functions
from . import layers
from . import cascade
from . import multikernel_layers
from . import utils
from . import specialized_kernels
from . import training

from .dsdgp import DSDGP
from .multitask_dsdgp import MultitaskDSDGP

__all__ = ["DSDGP", "MultitaskDSDGP", "layers", "cascade", "multikernel_layers", "utils", "specialized_kernels",
           "kernels", "likelihoods", "mean_functions", "settings", "training"]
