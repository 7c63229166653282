"""Mirror of the reference's nerf/mip_methods.py: maxBlurFilter (hot path) and the IPE ops."""
from . import ops


def maxBlurFilter(weights, alpha):
    """2-tap max then 2-tap mean + alpha (reference nerf/mip_methods.py:61-66)."""
    import torch
    if torch.is_grad_enabled() and weights.requires_grad:
        from .train_engine import MaxBlur
        return MaxBlur.apply(weights, alpha)
    return ops.max_blur(weights, alpha)


def ipe_feature(zvals, cam_rays, freq_lvs, r):
    """Integrated positional encoding of conical frustums (reference nerf/mip_methods.py:47-58).

    zvals (R, C+1), cam_rays (R, 6) -> (feature (R, C, 6L), mu (R, C, 3), mu_t (R, C)).
    Keeps the reference's batch-global direction norm (mip_methods.py:31).
    """
    return ops.ipe(zvals, cam_rays, freq_lvs, r)
