"""Mirror of the reference's nerf/ref_func.py (Ref-NeRF integrated directional encoding), forward only.

generate_ide_fn(deg_view) builds the (m, l) list and the spherical-harmonic coefficient matrix on the host exactly as
/root/reference/nerf/ref_func.py:38-76 does (float64 arithmetic, rounded to fp32 on assignment; `math.factorial`
instead of the removed `np.math.factorial`) and returns a closure that evaluates the encoding in the CUDA kernel
`ide_kernel` (nb2_ide).
"""
import math

import numpy as np
import torch

from . import ops


def generalized_binomial_coeff(a, k):
    return float(np.prod(a - np.arange(k))) / math.factorial(k)


def assoc_legendre_coeff(l, m, k):
    return ((-1) ** m * 2 ** l * math.factorial(l) / math.factorial(k) / math.factorial(l - k - m) *
            generalized_binomial_coeff(0.5 * (l + k + m - 1.0), l))


def sph_harm_coeff(l, m, k):
    return math.sqrt((2.0 * l + 1.0) * math.factorial(l - m) / (4.0 * math.pi * math.factorial(l + m))) * assoc_legendre_coeff(l, m, k)


def get_ml_array(deg_view):
    ml_list = []
    for i in range(deg_view):
        l = 2 ** i
        for m in range(l + 1):
            ml_list.append((m, l))
    return np.array(ml_list).T


def generate_ide_fn(deg_view):
    if deg_view > 5:
        raise ValueError("Only deg_view of at most 5 is numerically stable.")
    ml_array = get_ml_array(deg_view)
    l_max = 2 ** (deg_view - 1)
    mat = torch.zeros(l_max + 1, ml_array.shape[1])
    for i, (m, l) in enumerate(ml_array.T):
        for k in range(l - m + 1):
            mat[k, i] = sph_harm_coeff(int(l), int(m), k)
    ml = torch.from_numpy(ml_array.astype(np.int32)).contiguous()
    cache = {}

    def tables(dev):
        """(mat (n_pow, n_pairs) fp32, ml (2, n_pairs) int32) on `dev` (also read by the backward kernel)."""
        if dev not in cache:
            cache[dev] = (mat.to(dev), ml.to(dev))
        return cache[dev]

    def integrated_dir_enc_fn(xyz, kappa_inv):
        """xyz [..., 3] directions, kappa_inv [..., 1] -> [..., 2 * n_pairs] (real parts, then imaginary parts)."""
        return ops.ide(xyz, kappa_inv, *tables(xyz.device))

    integrated_dir_enc_fn.tables = tables
    return integrated_dir_enc_fn
