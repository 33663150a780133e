"""Numerical constants the reference reads from gpflow.settings (dgplib/utils.py:12-13, dgplib/dsdgp.py:178)."""
import torch

float_type = torch.float64      # gpflow.settings.float_type
jitter_level = 1e-6             # gpflow.settings.numerics.jitter_level
positive_lower = 1e-6           # gpflow.transforms.Log1pe lower bound:  theta = softplus(u) + 1e-6
# find_weights (dgplib/layers.py:97-121): "numpy" = np.linalg.svd on the host exactly as the reference (default, so that
# initial weights carry LAPACK's sign choice); "cuda" = dgp_pca_weights on the GPU (csrc/pca.cu; same vectors up to a
# per-vector sign, fixed by the convention in include/dgp_b200.h) for training sets where the host SVD is slow
pca_backend = "numpy"
