"""Numerical constants the reference reads from gpflow.settings (dgplib/utils.py:12-13, dgplib/dsdgp.py:178)."""
import torch

float_type = torch.float64      # gpflow.settings.float_type
jitter_level = 1e-6             # gpflow.settings.numerics.jitter_level
positive_lower = 1e-6           # gpflow.transforms.Log1pe lower bound:  theta = softplus(u) + 1e-6
