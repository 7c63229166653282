"""CPU: the ray-by-ray parity theorem (tests/parity_tools.py) exercised on a stand-in engine, so that the checker itself is
tested without a GPU.  The stand-in is the reference algorithm with the proposal densities evaluated in fp64 and rounded
to fp32 -- a perturbation of ~1e-6 relative, the same size as the tensor-core kernels'.  It shows the point of the
theorem: even the reference's own arithmetic, evaluated more precisely, moves some rays by more than 1e-4."""
import numpy as np
import torch

import nerf_b200
from oracle import nerf_oracle as O
from tests.parity_tools import assert_render_parity, draw_bounds, render_parity_report


def standin(H, W, perturb):
    pose = nerf_b200.pose_spherical(30.0, -30.0, 4.0)[:3, :]
    focal = nerf_b200.fov2Focal(0.6911112070083618, (H, W))[0]
    rays = O.generate_rays(pose, H, W, focal)
    sp, sn = O.make_params("proposal", 1, "smooth"), O.make_params("nerf", 2, "smooth")
    jit, u = O.det_uniform((H * W, 64), 31, 0.0, 1.0), O.det_uniform((H * W, 129), 32, 0.0, 1.0)
    base = torch.linspace(2.0, 6.0, 64)
    ref = O.render_rays(sp, sn, rays, base, jit, u, 2.0, 6.0, 128, white_bkg=True)
    pts = rays[:, None, :3] + ref["z_coarse"][..., None] * rays[:, None, 3:]
    sig = torch.from_numpy(O.np_forward("proposal", sp, pts.numpy()).astype(np.float32)).reshape(H * W, 64)
    sig = sig * (1.0 + perturb)
    w = O.max_blur(O.weights_from_sigma(sig, ref["z_coarse"], rays[:, 3:]), 0.01)
    zf, below = O.inverse_sample(w, ref["z_coarse"], u, sort=True)
    zk = zf[:, :-1]
    comp = O.composite(O.nerf_forward(sn, O.length2pts(rays, zk)), zk, rays[:, 3:], True, (2.0, 6.0))
    eng = dict(rgb=comp["rgb"], depth=comp["depth"], z_coarse=ref["z_coarse"], sigma_prop=sig, z_fine=zk, below_fine=below[:, :-1])
    return render_parity_report(O, sp, sn, rays, base, jit, u, 2.0, 6.0, eng)


def test_theorem_holds_for_an_ulp_level_perturbation():
    rep = standin(32, 32, 0.0)
    assert rep["B_rays"] > 0                      # flips do happen at this perturbation size ...
    assert_render_parity(rep, max_over_frac=0.02, max_flip_frac=0.05)  # ... and every one of them is within the reference's own bound


def test_theorem_rejects_a_real_regression():
    """A 1e-3 relative density error (what a dropped lo-term or a wrong bias would look like) must NOT pass."""
    rep = standin(24, 24, 1e-3)
    try:
        assert_render_parity(rep)
    except AssertionError:
        return
    raise AssertionError(f"a 1e-3 density regression passed the parity theorem: {rep}")


def test_draw_bounds_covers_actual_perturbations():
    g = torch.Generator().manual_seed(3)
    w = torch.rand(200, 62, generator=g) ** 8            # peaky weights: many near-empty bins
    z = torch.linspace(2.0, 6.0, 64).expand(200, 64) + torch.rand(200, 64, generator=g) * 0.03
    mids = 0.5 * (z[:, 1:] + z[:, :-1])
    u = torch.rand(200, 129, generator=g)
    cdf = O.build_cdf(w)
    noise = (torch.rand(200, 63, generator=g) - 0.5) * 2e-6
    noise[:, 0] = 0.0
    z0, b0, _ = O.invert_cdf(cdf, mids, u)
    z1, b1, _ = O.invert_cdf(cdf + noise, mids, u)
    bd, near = draw_bounds(cdf, mids, u, noise.abs().amax(-1, keepdim=True))
    assert bool(((z1 - z0).abs() <= bd).all())
    assert bool(((b0 == b1) | near).all())
