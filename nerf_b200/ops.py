"""Tensor-level wrappers over the C ABI (one function per entry point of include/nerf_b200.h).

Everything here takes and returns CUDA tensors on the current device and enqueues on the
current CUDA stream.  Reference-named functions (positional_encoding, inverseSample, ...) live
in the sibling modules that mirror the reference's package layout and call into this file.
"""
import ctypes

import torch

from . import _lib
from ._lib import NB2Error, PRECISIONS, check, f32, handle, load, ptr, stream_ptr

_default_precision = "fp16x3"


def set_default_precision(name):
    """'fp32' (CUDA cores, strict), 'fp16x3' (tcgen05, fp32-faithful hi/lo split; default), 'bf16' / 'fp16'
    (tcgen05 single pass) or 'bf16x3'."""
    global _default_precision
    if name not in PRECISIONS:
        raise ValueError(f"unknown precision {name!r}; choose from {sorted(PRECISIONS)}")
    _default_precision = name


def get_default_precision():
    return _default_precision


def _prec(name):
    name = _default_precision if name is None else name
    if name not in PRECISIONS:
        raise ValueError(f"unknown precision {name!r}; choose from {sorted(PRECISIONS)}")
    return PRECISIONS[name]


def _seed_from_torch():
    """A 63-bit Philox key drawn from torch's default CPU generator (so torch.manual_seed applies)."""
    return int(torch.randint(0, 2**62, (1,), dtype=torch.int64).item())


# ---- a1 ---------------------------------------------------------------------------------------
def generate_rays(pose, H, W, focal_x, focal_y, pix_offset=0, n_rays=None, out=None):
    pose = f32(pose[:3, :4])
    n = H * W - pix_offset if n_rays is None else n_rays
    rays = out if out is not None else torch.empty((n, 6), dtype=torch.float32, device=pose.device)
    check(load().nb2_generate_rays(handle(pose.device), ptr(pose), H, W, float(focal_x), float(focal_y), pix_offset, n,
                                   ptr(rays), stream_ptr(pose.device)))
    return rays


# ---- a3 + a4 ----------------------------------------------------------------------------------
def sample_coarse(rays, base_z, resolution, jitter=None, seed=0, ray_offset=0, want_pts=True):
    rays, base_z = f32(rays), f32(base_z)
    R, P = rays.shape[0], base_z.numel()
    if jitter is not None:
        jitter = f32(jitter).view(R, P)
    z = torch.empty((R, P), dtype=torch.float32, device=rays.device)
    pts = torch.empty((R, P, 3), dtype=torch.float32, device=rays.device) if want_pts else None
    check(load().nb2_sample_coarse(handle(rays.device), ptr(rays), ptr(base_z), ptr(jitter), float(resolution), seed,
                                   ray_offset, R, P, ptr(z), ptr(pts), stream_ptr(rays.device)))
    return z, pts


# ---- a5 ---------------------------------------------------------------------------------------
def posenc(x, levels):
    x = f32(x)
    dims = x.shape[-1]
    n = x.numel() // dims
    out = torch.empty((n, 2 * dims * levels), dtype=torch.float32, device=x.device)
    check(load().nb2_posenc(handle(x.device), ptr(x), n, dims, levels, ptr(out), stream_ptr(x.device)))
    return out


# ---- a14 --------------------------------------------------------------------------------------
def ipe(zvals, cam_rays, levels, radius):
    zvals, cam_rays = f32(zvals), f32(cam_rays)
    R, C = zvals.shape[0], zvals.shape[1] - 1
    dev = zvals.device
    feat = torch.empty((R, C, 6 * levels), dtype=torch.float32, device=dev)
    mu = torch.empty((R, C, 3), dtype=torch.float32, device=dev)
    mu_t = torch.empty((R, C), dtype=torch.float32, device=dev)
    scratch = torch.empty(2, dtype=torch.float64, device=dev)
    check(load().nb2_ipe(handle(dev), ptr(zvals), ptr(cam_rays), R, C, levels, float(radius), ptr(feat), ptr(mu),
                         ptr(mu_t), ptr(scratch), stream_ptr(dev)))
    return feat, mu, mu_t


# ---- a7 ---------------------------------------------------------------------------------------
_ACTS = {"relu": 0, "softplus": 1, "identity": 2}


def weights_from_sigma(sigma, z, dirs=None, act="relu"):
    sigma, z = f32(sigma), f32(z)
    R, P = z.shape
    stride = 0
    if dirs is not None:
        dirs = f32(dirs)
        stride = dirs.shape[-1]
    w = torch.empty((R, P), dtype=torch.float32, device=z.device)
    check(load().nb2_weights_from_sigma(handle(z.device), ptr(sigma), ptr(z), ptr(dirs), stride, R, P, _ACTS[act],
                                        ptr(w), stream_ptr(z.device)))
    return w


# ---- a8 ---------------------------------------------------------------------------------------
def max_blur(weights, alpha):
    weights = f32(weights)
    P = weights.shape[-1]
    R = weights.numel() // P
    out = torch.empty_like(weights)
    check(load().nb2_max_blur(handle(weights.device), ptr(weights), R, P, float(alpha), ptr(out), stream_ptr(weights.device)))
    return out


# ---- a9 ---------------------------------------------------------------------------------------
def sample_pdf(bins, weights, n_draw, u=None, seed=None, ray_offset=0):
    bins, weights = f32(bins), f32(weights)
    R, B = bins.shape
    if weights.shape != (R, B - 1):
        raise NB2Error(f"sample_pdf: weights must be (R, len(bins) - 1), got {tuple(weights.shape)} for bins {tuple(bins.shape)}")
    dev = bins.device
    if u is not None:
        u = f32(u).view(R, n_draw)
    elif seed is None:
        seed = _seed_from_torch()
    samples = torch.empty((R, n_draw), dtype=torch.float32, device=dev)
    below = torch.empty((R, n_draw), dtype=torch.int64, device=dev)
    above = torch.empty((R, n_draw), dtype=torch.int64, device=dev)
    check(load().nb2_sample_pdf(handle(dev), ptr(bins), ptr(weights), ptr(u), seed or 0, ray_offset, R, B, n_draw,
                                ptr(samples), ptr(below), ptr(above), stream_ptr(dev)))
    return samples, below, above


def inverse_sample(weights, z, n_draw, sort=False, u=None, seed=None, ray_offset=0):
    weights, z = f32(weights), f32(z)
    R, P = z.shape
    dev = z.device
    if u is not None:
        u = f32(u).view(R, n_draw)
    elif seed is None:
        seed = _seed_from_torch()
    samples = torch.empty((R, n_draw), dtype=torch.float32, device=dev)
    below = torch.empty((R, n_draw), dtype=torch.int64, device=dev)
    check(load().nb2_inverse_sample(handle(dev), ptr(weights), ptr(z), ptr(u), seed or 0, ray_offset, R, P, n_draw,
                                    1 if sort else 0, ptr(samples), ptr(below), stream_ptr(dev)))
    return samples, below


def search_cdf(cdf, u):
    cdf, u = f32(cdf), f32(u)
    R, B = cdf.shape
    N = u.shape[1]
    inds = torch.empty((R, N), dtype=torch.int64, device=cdf.device)
    check(load().nb2_search_cdf(handle(cdf.device), ptr(cdf), ptr(u), R, B, N, ptr(inds), stream_ptr(cdf.device)))
    return inds


def resample(sigma, z, rays, n_draw, blur_alpha=0.01, u=None, seed=0, ray_offset=0, softplus=False, want_below=False):
    sigma, z, rays = f32(sigma), f32(z), f32(rays)
    R, P = z.shape
    if u is not None:
        u = f32(u).view(R, n_draw)
    out = torch.empty((R, n_draw - 1), dtype=torch.float32, device=z.device)
    below = torch.empty((R, n_draw - 1), dtype=torch.int64, device=z.device) if want_below else None
    flags = _lib.DENSITY_SOFTPLUS if softplus else 0
    check(load().nb2_resample(handle(z.device), ptr(sigma), ptr(z), ptr(rays), ptr(u), seed, ray_offset, R, P, n_draw,
                              float(blur_alpha), flags, ptr(out), ptr(below), stream_ptr(z.device)))
    return (out, below) if want_below else out


# ---- a10 / a13 ----------------------------------------------------------------------------------
def length2pts(rays, z):
    rays, z = f32(rays), f32(z)
    R, P = z.shape
    pts = torch.empty((R, P, 6), dtype=torch.float32, device=z.device)
    check(load().nb2_length2pts(handle(z.device), ptr(rays), ptr(z), R, P, ptr(pts), stream_ptr(z.device)))
    return pts


def coarse_fine_merge(rays, c_z, f_z, f_inds=None):
    """-> (pts, z) or, with f_inds (R,F) int64, (pts, z, all_inds (R,F+C), sort_inds (R,F+C-1))  (nerf_base.py:58-73)."""
    rays, c_z, f_z = f32(rays), f32(c_z), f32(f_z)
    R, C, F = c_z.shape[0], c_z.shape[1], f_z.shape[1]
    dev = c_z.device
    z = torch.empty((R, C + F - 1), dtype=torch.float32, device=dev)
    pts = torch.empty((R, C + F - 1, 6), dtype=torch.float32, device=dev)
    if f_inds is None:
        check(load().nb2_coarse_fine_merge(handle(dev), ptr(rays), ptr(c_z), ptr(f_z), R, C, F, ptr(z), ptr(pts), stream_ptr(dev)))
        return pts, z
    f_inds = f_inds.to(torch.int64).contiguous()
    all_inds = torch.empty((R, C + F), dtype=torch.int64, device=dev)
    sort_inds = torch.empty((R, C + F - 1), dtype=torch.int64, device=dev)
    check(load().nb2_coarse_fine_merge_inds(handle(dev), ptr(rays), ptr(c_z), ptr(f_z), ptr(f_inds), R, C, F, ptr(z), ptr(pts), ptr(all_inds),
                                            ptr(sort_inds), stream_ptr(dev)))
    return pts, z, all_inds, sort_inds


# ---- f1: training-side callers (forward only) -------------------------------------------------------
_BASE_Z = {}


def _base_z(start, end, n, dev):
    """The reference builds torch.linspace on the host every step (utils.py:88); the same values, uploaded once: a pageable
    host-to-device copy waits for the stream, which would serialise the host with every training step."""
    key = (start, end, n, dev)
    z = _BASE_Z.get(key)
    if z is None:
        if len(_BASE_Z) > 64:
            _BASE_Z.clear()
        z = _BASE_Z[key] = torch.linspace(start, end, n).to(dev)
    return z


def valid_sampler(rgbs, coords, cam_tf, ray_num, point_num, focal_x, focal_y, near, far, indices=None, jitter=None, seed=None):
    rgbs, cam_tf = f32(rgbs), f32(cam_tf[:3, :4])
    coords = coords.to(torch.int64).contiguous()
    dev = rgbs.device
    resolution = (far - near) / point_num
    base_z = _base_z(float(near), float(far - resolution), int(point_num), dev)     # utils.py:88
    if indices is not None:
        indices = indices.to(device=dev, dtype=torch.int64).contiguous()
    if jitter is not None:
        jitter = f32(jitter).view(ray_num, point_num)
    if seed is None:
        seed = _seed_from_torch() if (indices is None or jitter is None) else 0
    pts = torch.empty((ray_num, point_num, 3), dtype=torch.float32, device=dev)
    lengths = torch.empty((ray_num, point_num), dtype=torch.float32, device=dev)
    rgb = torch.empty((ray_num, 3), dtype=torch.float32, device=dev)
    rays = torch.empty((ray_num, 6), dtype=torch.float32, device=dev)
    check(load().nb2_valid_sampler(handle(dev), ptr(rgbs), ptr(coords), ptr(cam_tf), ptr(indices), ptr(base_z), ptr(jitter),
                                   float(focal_x), float(focal_y), float(resolution), seed, 0, coords.shape[0], ray_num, point_num,
                                   ptr(pts), ptr(lengths), ptr(rgb), ptr(rays), stream_ptr(dev)))
    return pts, lengths, rgb, rays


def get_bounds(weights, inds):
    weights = f32(weights)
    inds = inds.to(torch.int64).contiguous()
    R, P = weights.shape
    K = inds.shape[1]
    out = torch.empty((R, K - 1), dtype=torch.float32, device=weights.device)
    check(load().nb2_get_bounds(handle(weights.device), ptr(weights), ptr(inds), R, P, K, ptr(out), stream_ptr(weights.device)))
    return out


# ---- weights / MLP --------------------------------------------------------------------------------
def net_create(kind, device):
    """A packed-network slot of `kind` on `device`'s handle (include/nerf_b200.h: nb2_net_create)."""
    out = ctypes.c_int(-1)
    check(load().nb2_net_create(handle(device), kind, ctypes.byref(out)))
    _NET_KIND[(torch.device(device).index, int(out.value))] = kind
    return int(out.value)


_NET_KIND = {}


def net_kind(net_id, device):
    """Kind (NET_PROPOSAL | NET_NERF) of a slot; the default slots 0 / 1 are their own kinds."""
    return net_id if net_id < 2 else _NET_KIND[(torch.device(device).index, net_id)]


def net_destroy(net_id, device):
    check(load().nb2_net_destroy(handle(device), net_id))
    _NET_KIND.pop((torch.device(device).index, net_id), None)


def pack_weights(net_id, weights, biases, pos_levels, dir_levels, hidden, device=None):
    """weights/biases: lists of fp32 CUDA tensors in reference state_dict order; net_id: a slot id."""
    ws = [f32(w.detach()) for w in weights]
    bs = [f32(b.detach()) for b in biases]
    n = len(ws)
    W = (ctypes.c_void_p * n)(*[w.data_ptr() for w in ws])
    B = (ctypes.c_void_p * n)(*[b.data_ptr() for b in bs])
    dev = ws[0].device if device is None else device
    check(load().nb2_pack_weights(handle(dev), net_id, W, B, n, pos_levels, dir_levels, hidden, stream_ptr(dev)))
    return ws, bs  # keep-alive until the stream has consumed them


def mlp_forward(net_id, pts, precision=None):
    pts = f32(pts)
    stride = pts.shape[-1]
    n = pts.numel() // stride
    dev = pts.device
    out = torch.empty((n, 4) if net_kind(net_id, dev) == _lib.NET_NERF else (n,), dtype=torch.float32, device=dev)
    check(load().nb2_mlp_forward(handle(dev), net_id, _prec(precision), ptr(pts), stride, n, ptr(out), stream_ptr(dev)))
    return out


def mlp_forward_encoded(net_id, pts, encoded, precision=None):
    """Proposal network on externally encoded position features (n, 6 * levels) + raw points (n, >=3)."""
    pts, encoded = f32(pts), f32(encoded)
    stride = pts.shape[-1]
    n = pts.numel() // stride
    if encoded.dim() < 2 or encoded.numel() // encoded.shape[-1] != n:
        raise NB2Error(f"mlp_forward_encoded: encoded features must hold one row per point ({n}), got shape {tuple(encoded.shape)}")
    out = torch.empty((n,), dtype=torch.float32, device=pts.device)
    check(load().nb2_mlp_forward_encoded(handle(pts.device), net_id, _prec(precision), ptr(pts), stride, ptr(encoded), n, ptr(out), stream_ptr(pts.device)))
    return out


# ---- a12 ------------------------------------------------------------------------------------------
def composite(rgbo, z, dirs, white_bkg=False, near_far=None, want_weights=True, aux=None):
    """-> (rgb, weights, depth, acc) [+ aux_out (R) = sum_i w_i aux_i when aux (R,P) is given]."""
    rgbo, z, dirs = f32(rgbo), f32(z), f32(dirs)
    R, P = z.shape
    dev = z.device
    rgb = torch.empty((R, 3), dtype=torch.float32, device=dev)
    w = torch.empty((R, P), dtype=torch.float32, device=dev) if want_weights else None
    depth = torch.empty((R,), dtype=torch.float32, device=dev) if near_far is not None else None
    acc = torch.empty((R,), dtype=torch.float32, device=dev)
    near, far = near_far if near_far is not None else (0.0, 1.0)
    if aux is not None:
        aux = f32(aux).view(R, P)
        aux_out = torch.empty((R,), dtype=torch.float32, device=dev)
        check(load().nb2_composite_aux(handle(dev), ptr(rgbo), ptr(z), ptr(dirs), dirs.shape[-1], R, P, _lib.WHITE_BKG if white_bkg else 0,
                                       float(near), float(far), ptr(aux), ptr(rgb), ptr(w), ptr(depth), ptr(acc), ptr(aux_out), stream_ptr(dev)))
        return rgb, w, depth, acc, aux_out
    check(load().nb2_composite(handle(dev), ptr(rgbo), ptr(z), ptr(dirs), dirs.shape[-1], R, P,
                               _lib.WHITE_BKG if white_bkg else 0, float(near), float(far), ptr(rgb), ptr(w), ptr(depth),
                               ptr(acc), stream_ptr(dev)))
    return rgb, w, depth, acc


def dot3(a, b):
    """a (..., 3) . b (3) -> (...)   (normal @ cam_dir, nerf_base.py:111)."""
    a, b = f32(a), f32(b).reshape(3)
    out = torch.empty(a.shape[:-1], dtype=torch.float32, device=a.device)
    check(load().nb2_dot3(handle(a.device), ptr(a), ptr(b), a.numel() // 3, ptr(out), stream_ptr(a.device)))
    return out


# ---- next rows: Ref-NeRF forward helpers ------------------------------------------------------------
def ide(xyz, kappa_inv, mat, ml):
    """Integrated directional encoding (ref_func.py:78-108).  mat (n_pow, n_pairs) fp32, ml (2, n_pairs) int32 on the device."""
    xyz, kappa_inv = f32(xyz), f32(kappa_inv)
    n = xyz.numel() // 3
    if kappa_inv.numel() != n:
        raise NB2Error(f"ide: kappa_inv must hold one value per direction ({kappa_inv.numel()} for {n} directions)")
    n_pow, n_pairs = mat.shape
    out = torch.empty((*xyz.shape[:-1], 2 * n_pairs), dtype=torch.float32, device=xyz.device)
    check(load().nb2_ide(handle(xyz.device), ptr(xyz), ptr(kappa_inv), n, ptr(mat), ptr(ml), n_pairs, n_pow, ptr(out), stream_ptr(xyz.device)))
    return out


def linear_to_srgb(linear):
    linear = f32(linear)
    out = torch.empty_like(linear)
    check(load().nb2_linear_to_srgb(handle(linear.device), ptr(linear), linear.numel(), ptr(out), stream_ptr(linear.device)))
    return out


# ---- the fused path ---------------------------------------------------------------------------------
def render_rays(rays, base_z, near, far, n_fine=128, white_bkg=False, precision=None, jitter=None, u=None, seed=0,
                ray_offset=0, resolution=None, blur_alpha=0.01, softplus=False, debug=False, workspace=None,
                prop_net_id=_lib.NET_PROPOSAL, nerf_net_id=_lib.NET_NERF, out=None, peer_rgb=(), ipe_radius=None):
    """rays (R,6) -> dict(rgb (R,3), depth (R), acc (R) [, z_coarse, sigma_prop, z_fine, below_fine]).

    `out` re-uses a previous result's buffers (no allocation in the step); `peer_rgb`: device pointers (ints) of the
    other GPUs' full image buffers, which receive this shard's rgb rows at row offset `ray_offset` straight from the
    compositing epilogue (nb2_render_params.peer_rgb)."""
    rays, base_z = f32(rays), f32(base_z)
    R, Pc = rays.shape[0], base_z.numel()
    dev = rays.device
    p = _lib.RenderParams()
    p.n_coarse, p.n_fine = Pc, n_fine
    p.near_t, p.far_t = float(near), float(far)
    p.resolution = float((far - near) / n_fine if resolution is None else resolution)
    p.blur_alpha = float(blur_alpha)
    p.flags = (_lib.WHITE_BKG if white_bkg else 0) | (_lib.DENSITY_SOFTPLUS if softplus else 0) | (_lib.PROPOSAL_IPE if ipe_radius else 0)
    p.ipe_radius = float(ipe_radius or 0.0)
    p.precision = _prec(precision)
    p.seed, p.ray_offset = seed, ray_offset
    p.prop_net_id, p.nerf_net_id = prop_net_id, nerf_net_id
    p.n_peers = len(peer_rgb)
    for q, pp in enumerate(peer_rgb):
        p.peer_rgb[q] = pp
    lib = load()
    need = lib.nb2_render_workspace_bytes(R, ctypes.byref(p))
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
    if jitter is not None:
        jitter = f32(jitter).view(R, Pc)
    if u is not None:
        u = f32(u).view(R, n_fine + 1)
    if out is None or out["rgb"].shape[0] != R or out["rgb"].device != dev:
        out = {
            "rgb": torch.empty((R, 3), dtype=torch.float32, device=dev),
            "depth": torch.empty((R,), dtype=torch.float32, device=dev),
            "acc": torch.empty((R,), dtype=torch.float32, device=dev),
        }
    zc = sp = zf = bf = None
    if debug:
        zc = torch.empty((R, Pc), dtype=torch.float32, device=dev)
        sp = torch.empty((R, Pc), dtype=torch.float32, device=dev)
        zf = torch.empty((R, n_fine), dtype=torch.float32, device=dev)
        bf = torch.empty((R, n_fine), dtype=torch.int64, device=dev)
        out.update(z_coarse=zc, sigma_prop=sp, z_fine=zf, below_fine=bf)
    check(lib.nb2_render_rays(handle(dev), ctypes.byref(p), ptr(rays), ptr(base_z), ptr(jitter), ptr(u), R, ptr(out["rgb"]),
                              ptr(out["depth"]), ptr(out["acc"]), ptr(zc), ptr(sp), ptr(zf), ptr(bf), ptr(workspace),
                              workspace.numel(), stream_ptr(dev)))
    out["_workspace"] = workspace
    return out


def selftest_umma(A, B):
    """D = A @ B.T through one tcgen05 MMA sequence; A, B: (128, 64) bf16 CUDA tensors."""
    A = A.to(torch.bfloat16).contiguous()
    B = B.to(torch.bfloat16).contiguous()
    D = torch.empty((128, 128), dtype=torch.float32, device=A.device)
    scratch = torch.empty(16384, dtype=torch.uint8, device=A.device)
    check(load().nb2_selftest_umma(handle(A.device), ptr(A), ptr(B), ptr(scratch), ptr(D), stream_ptr(A.device)))
    return D
