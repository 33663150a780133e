"""dgplib/utils.py surface.  On the fused path neither function is called: tiling is index arithmetic inside the
kernels and sampling is their epilogue; they are kept for users of the reference API (torch tensors in / out)."""
import torch

from . import settings

jitter = settings.jitter_level
float_type = settings.float_type


def tile_over_samples(X, S):
    """[..] -> [S, ..] (dgplib/utils.py:15-18) as a stride-0 view: no copies."""
    return X.unsqueeze(0).expand(S, *X.shape)


def shape_as_list(X):
    return list(X.shape)


def normal_sample(mean, var, full_cov=False, eps=None):
    """mean + z * var ** 0.5 (dgplib/utils.py:25-27); `eps` injects the draw for reproducibility."""
    if full_cov:
        raise NotImplementedError("full-covariance sampling is outside the B200 hot path (SURVEY.md 8f, rank 3)")
    z = torch.randn_like(mean) if eps is None else eps
    return mean + z * var ** 0.5
