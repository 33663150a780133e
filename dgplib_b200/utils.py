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
        # mean [S,N,D], var [S,N,N,D], z [S,D,N]: chol(var + jitter I) z per (sample, output), dgplib/utils.py:28-36
        from .engine import Engine  # noqa: F401  (CUDA path)
        from . import _cabi
        _cabi.require_cuda()
        S, N, D = mean.shape
        z = torch.randn(S, D, N, dtype=mean.dtype, device=mean.device) if eps is None else eps
        eng = _sampler()
        out = eng.chol_sample(var.permute(0, 3, 1, 2).reshape(S * D, N, N).contiguous(), z.reshape(S * D, N, 1),
                              mean.permute(0, 2, 1).reshape(S * D, N).contiguous())
        return out.reshape(S, D, N).permute(0, 2, 1).contiguous()
    z = torch.randn_like(mean) if eps is None else eps
    return mean + z * var ** 0.5


_SAMPLER = None


def _sampler():
    """A layer-less Engine: only its chol_sample (dgp_chol_sample) is used."""
    global _SAMPLER
    if _SAMPLER is None:
        from .engine import Engine
        _SAMPLER = Engine([], None)
    return _SAMPLER
