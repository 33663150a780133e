"""MultitaskDSDGP (dgplib/multitask_dsdgp.py): the task-id column of the input is re-appended to every layer's
sample (multitask_dsdgp.py:20).  The forward kernel writes that column directly (dgp_cond_forward with ldf = R + 1)."""
from .dsdgp import DSDGP


class MultitaskDSDGP(DSDGP):
    multitask = True
