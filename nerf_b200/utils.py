"""Mirror of the reference's nerf/utils.py for the hot path: inverse-CDF sampling and pose math."""
from collections.abc import Iterable

import numpy as np
import torch

from . import ops


def sample_pdf(bins, weights, N_samples, u=None):
    """Inverse-transform sampling (reference nerf/utils.py:108-133).

    bins (R, B), weights (R, B-1) -> (samples (R, N) fp32, below (R, N) int64, above (R, N) int64).
    `u` (R, N) injects the uniforms; by default they are drawn on the device with Philox keyed from
    torch's CPU generator (the reference draws torch.rand on the CPU and copies it over).
    """
    return ops.sample_pdf(bins, weights, N_samples, u=u)


def inverseSample(weights, coarse_depth, sample_pnum, sort=False, u=None):
    """reference nerf/utils.py:34-44.  Returns (z, below) when sort=True, z otherwise."""
    weights = weights.detach()
    z, below = ops.inverse_sample(weights, coarse_depth, sample_pnum, sort=sort, u=u)
    if sort:
        return z, below
    return z


def validSampler(rgbs, coords, cam_tf, ray_num, point_num, focal, near, far, output_samples=True, indices=None, jitter=None):
    """Training-side ray sampler (reference nerf/utils.py:72-94): random pixels of one image -> rays, true stratified
    depths and sample points.  `indices` / `jitter` inject the reference's CPU draws (torch.randint / torch.rand)."""
    if isinstance(focal, Iterable):
        fx, fy = float(focal[1]), float(focal[0])
    else:
        fx = fy = float(focal)
    pts, lengths, rgb, rays = ops.valid_sampler(rgbs, coords, cam_tf, ray_num, point_num, fx, fy, near, far, indices=indices, jitter=jitter)
    if output_samples:
        return pts, lengths, rgb, rays
    return rgb, rays


def fov2Focal(fov, img_size):
    """reference nerf/utils.py:96-105 (including the missing 1/2 for a scalar fov)."""
    if isinstance(fov, Iterable):
        if not isinstance(img_size, Iterable):
            raise ValueError("Error: If fov is iterable, img size should be iterable too, while we have typeof(img_size) =", type(img_size))
        return (0.5 * img_size[0] / np.tan(.5 * fov[1]), 0.5 * img_size[1] / np.tan(.5 * fov[0]))
    if img_size[0] == img_size[1]:
        img_size = img_size[0]
    focal = img_size / np.tan(.5 * fov)
    return (focal, focal)


def pose_spherical(theta, phi, radius):
    """Orbit camera pose (reference nerf/utils.py:136-159); 4x4 camera-to-world, CPU tensor."""
    th, ph = theta / 180. * np.pi, phi / 180. * np.pi
    trans = torch.tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, radius], [0, 0, 0, 1]], dtype=torch.float32)
    rot_phi = torch.tensor([[1, 0, 0, 0], [0, np.cos(ph), -np.sin(ph), 0], [0, np.sin(ph), np.cos(ph), 0], [0, 0, 0, 1]],
                           dtype=torch.float32)
    rot_theta = torch.tensor([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]],
                             dtype=torch.float32)
    swap = torch.tensor([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=torch.float32)
    return swap @ (rot_theta @ (rot_phi @ trans))
