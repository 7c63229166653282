"""oracle/nerf_oracle.py — CPU restatement of the reference's ray-marching path.

TEST INFRASTRUCTURE ONLY.  Nothing under nerf_b200/ imports this file; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may.  It is the
checker, never the thing measured or shipped.

What it is: the algorithm of Enigmatisms/NeRF's `render_image` hot path (SURVEY.md §8a rows
a1-a14) written as plain functions over explicit parameter dictionaries, with every random
draw passed in as an argument (the reference draws `torch.rand` on the CPU in the middle of the
computation: nerf/procedures.py:65, nerf/utils.py:115).  The arithmetic is fp32 PyTorch tensor
ops, because that is what the reference's arithmetic *is*: all its numerics live in the
third-party dependency PyTorch (requirements.txt pins torch==2.6.0; this image has 2.11.0 —
version drift that no reference test pins).  The functions are device-agnostic: tests run them
on the CPU at small sizes and on `cuda` for full-size parity (the reference's own GPU path).

Pinning: the reference ships no tests, golden vectors or fixtures.  The oracle
is pinned instead against outputs of the reference itself: tests/golden/make_golden.py imports
/root/reference/nerf unmodified (with import shims), runs its functions on seeded inputs and
commits the results as tests/golden/*.npz; tests/test_oracle_golden.py checks this file against
them.

Each function cites the reference lines it restates (paths relative to /root/reference).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# deterministic parameters / uniforms: nerf_b200/synthetic.py (data generation shared with bench.py; re-exported here)
# ----------------------------------------------------------------------------------------------
from nerf_b200.synthetic import (NERF_KEYS, PROPOSAL_KEYS, _mix64, det_state_dict, det_uniform, layer_shapes,  # noqa: E402,F401
                                 make_params, np_forward, params_to)


# ----------------------------------------------------------------------------------------------
# a5  positional encoding                                      nerf/nerf_helper.py:38-48
# ----------------------------------------------------------------------------------------------
def positional_encoding(x, levels):
    parts = []
    for f in range(levels):
        parts.append(torch.sin((2.0 ** f) * x))
        parts.append(torch.cos((2.0 ** f) * x))
    return torch.cat(parts, dim=-1)


# ----------------------------------------------------------------------------------------------
# a6  proposal MLP                                             nerf/addtional.py:61-72,88-96
# ----------------------------------------------------------------------------------------------
def _relu(x, masks, i):
    """ReLU, or -- when `masks` is given -- multiplication by a fixed 0/1 pattern: the gradient of a ReLU network is
    discontinuous in the sign of every pre-activation, so a gradient check against another implementation (whose forward
    pass differs by rounding) has to hold the activation pattern fixed to compare like with like."""
    return F.relu(x) if masks is None else x * masks[i].to(x.dtype)


def proposal_forward(sd, pts, pos_levels=10, encoded=None, relu_masks=None):
    """pts (..., 3) -> raw density (...).  `encoded` (..., 6L) replaces the sinusoidal features (addtional.py:89-91)."""
    enc = positional_encoding(pts, pos_levels) if encoded is None else encoded.reshape(*pts.shape[:-1], 6 * pos_levels)
    h = torch.cat((pts, enc), dim=-1)       # cat_origin
    for i, key in enumerate(PROPOSAL_KEYS[:-1]):
        h = _relu(F.linear(h, sd[key + ".weight"], sd[key + ".bias"]), relu_masks, i)
    return F.linear(h, sd["layers.8.weight"], sd["layers.8.bias"]).squeeze(-1)


# ----------------------------------------------------------------------------------------------
# a11  NeRF MLP                                                nerf/mip_model.py:41-60
# ----------------------------------------------------------------------------------------------
def nerf_forward(sd, pts6, pos_levels=10, dir_levels=4, relu_masks=None):
    """pts6 (..., 6) = [xyz, dir] -> (..., 4) = [sigmoid rgb, raw sigma].  relu_masks: see _relu (8 patterns: 7 trunk layers + rgb_layer.0)."""
    x, d = pts6[..., :3], pts6[..., 3:6]
    rot = d / d.norm(dim=-1, keepdim=True)                                    # :44-45
    enc_x = torch.cat((x, positional_encoding(x, pos_levels)), dim=-1)       # :50-51
    enc_r = torch.cat((rot, positional_encoding(rot, dir_levels)), dim=-1)   # :52
    h = enc_x
    for i, key in enumerate(NERF_KEYS[:4]):                                   # lin_block1  :54
        h = _relu(F.linear(h, sd[key + ".weight"], sd[key + ".bias"]), relu_masks, i)
    h = torch.cat((enc_x, h), dim=-1)                                         # skip concat :55
    for i, key in enumerate(NERF_KEYS[4:7]):                                  # lin_block2  :56
        h = _relu(F.linear(h, sd[key + ".weight"], sd[key + ".bias"]), relu_masks, 4 + i)
    sigma = F.linear(h, sd["opacity_head.0.weight"], sd["opacity_head.0.bias"])   # :57
    b = F.linear(h, sd["bottle_neck.0.weight"], sd["bottle_neck.0.bias"])         # :58
    t = _relu(F.linear(torch.cat((b, enc_r), dim=-1), sd["rgb_layer.0.weight"], sd["rgb_layer.0.bias"]), relu_masks, 7)
    rgb = torch.sigmoid(F.linear(t, sd["rgb_layer.2.weight"], sd["rgb_layer.2.bias"]))  # :59
    return torch.cat((rgb, sigma), dim=-1)                                    # :60


# ----------------------------------------------------------------------------------------------
# a7  density -> weights                nerf/addtional.py:99-107 (relu) == nerf/nerf_base.py:79-86
# ----------------------------------------------------------------------------------------------
def weights_from_sigma(sigma, z, dirs=None, act=F.relu):
    if dirs is not None:
        z = z * dirs.norm(dim=-1, keepdim=True)
    big = torch.full((z.shape[0], 1), 1e10, dtype=z.dtype, device=z.device)
    delta = torch.cat((z[:, 1:] - z[:, :-1], big), dim=-1)
    mult = torch.exp(-(act(sigma) if act is not None else sigma) * delta)
    alpha = 1.0 - mult
    ones = torch.ones((z.shape[0], 1), dtype=z.dtype, device=z.device)
    trans = torch.cumprod(torch.cat((ones, mult + 1e-10), dim=-1), dim=-1)[:, :-1]
    return alpha * trans


# ----------------------------------------------------------------------------------------------
# a8  max-blur filter                                          nerf/mip_methods.py:61-66
# ----------------------------------------------------------------------------------------------
def max_blur(w, alpha):
    mx = torch.maximum(w[..., :-1], w[..., 1:])
    front = torch.cat((w[..., :1], mx), dim=-1)
    rear = torch.cat((mx, w[..., -1:]), dim=-1)
    return 0.5 * (front + rear) + alpha


# ----------------------------------------------------------------------------------------------
# a9  inverse-CDF sampling                                     nerf/utils.py:108-133, :34-44
# ----------------------------------------------------------------------------------------------
def build_cdf(weights):
    """pdf / cdf of sample_pdf (utils.py:110-113), in the reduction order the CUDA kernel documents:
    total = fp32(sum in fp64), cdf_k = fp32(running fp64 sum of fp32 pdf) — the latter is what
    torch.cumsum does on the CPU for fp32 input."""
    w = weights + 1e-5
    total = w.double().sum(dim=-1, keepdim=True).to(w.dtype)
    pdf = w / total
    cdf = torch.cumsum(pdf.double(), dim=-1).to(w.dtype)
    return torch.cat((torch.zeros_like(cdf[..., :1]), cdf), dim=-1)


def build_cdf_torch(weights):
    """The same with torch's own fp32 `sum` (utils.py:110-113 verbatim semantics; reduction order is
    whatever ATen picks on this machine)."""
    w = weights + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    return torch.cat((torch.zeros_like(cdf[..., :1]), cdf), dim=-1)


def invert_cdf(cdf, bins, u):
    """utils.py:119-131: searchsorted(right=True), clamp, gather, guarded lerp."""
    u = u.contiguous()
    inds = torch.searchsorted(cdf.contiguous(), u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_b, cdf_a = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    bins_b, bins_a = torch.gather(bins, -1, below), torch.gather(bins, -1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_b) / denom
    return bins_b + t * (bins_a - bins_b), below, above


def sample_pdf(bins, weights, u, torch_sum=False):
    cdf = build_cdf_torch(weights) if torch_sum else build_cdf(weights)
    return invert_cdf(cdf, bins, u)


def inverse_sample(weights, z, u, sort=True, torch_sum=False):
    """utils.py:34-44: mid-point bins, interior weights, optional sort with index gather."""
    mids = 0.5 * (z[..., 1:] + z[..., :-1])
    samples, below, _ = sample_pdf(mids, weights[..., 1:-1], u, torch_sum=torch_sum)
    if sort:
        samples, order = torch.sort(samples, dim=-1)
        below = torch.gather(below, -1, order)
    return samples, below


# ----------------------------------------------------------------------------------------------
# a10 / a13  sample points                                     nerf/nerf_base.py:52-56, :58-73
# ----------------------------------------------------------------------------------------------
def length2pts(rays, z):
    pts = rays[:, None, :3] + rays[:, None, 3:] * z[:, :, None]
    return torch.cat((pts, rays[:, None, 3:].expand(-1, z.shape[1], -1)), dim=-1)


def coarse_fine_merge(rays, c_z, f_z):
    z, _ = torch.sort(torch.cat((f_z, c_z), dim=-1), dim=-1)
    z = z[..., :-1]
    return length2pts(rays, z), z


# ----------------------------------------------------------------------------------------------
# a12  alpha compositing                                       nerf/nerf_base.py:90-113
# ----------------------------------------------------------------------------------------------
def composite(rgbo, z, dirs, white_bkg=False, near_far=None):
    depth = z * dirs.norm(dim=-1, keepdim=True)
    w = weights_from_sigma(rgbo[..., -1], depth, None, F.relu)
    rgb = torch.sum(w[:, :, None] * rgbo[..., :3], dim=-2)
    acc = torch.sum(w, -1)
    if white_bkg:
        rgb = rgb + (1.0 - acc[..., None])
    out = {"rgb": rgb, "weights": w, "acc": acc}
    if near_far is not None:
        near, far = near_far
        out["depth"] = (torch.sum(w * depth, dim=-1) - near) / (far - near)
    return out


# ----------------------------------------------------------------------------------------------
# a1  ray generation                                           nerf/procedures.py:43-51
# ----------------------------------------------------------------------------------------------
def generate_rays(pose, H, W, focal):
    """pose (3,4) -> rays (H*W, 6) in raster order: [origin, R @ (cx/fx, cy/fy, -1)] (un-normalised)."""
    col, row = torch.meshgrid(torch.arange(W), torch.arange(H), indexing="xy")
    coords = torch.stack((col - W / 2, H / 2 - row), dim=-1).to(pose.device) + 0.5
    fx, fy = (focal[1], focal[0]) if isinstance(focal, (tuple, list)) else (focal, focal)
    coords = torch.stack((coords[..., 0] / fx, coords[..., 1] / fy), dim=-1)
    coords = torch.cat((coords, -torch.ones(H, W, 1, dtype=torch.float32, device=pose.device)), dim=-1)
    dirs = torch.sum(coords.unsqueeze(-2) * pose[..., :-1], dim=-1)
    origin = pose[:, -1].expand(H, W, -1)
    return torch.cat((origin, dirs), dim=-1).reshape(-1, 6)


# ----------------------------------------------------------------------------------------------
# a14  integrated positional encoding                          nerf/mip_methods.py:15-58
# ----------------------------------------------------------------------------------------------
def ipe_feature(zvals, rays, levels, r):
    mid = (zvals[:, 1:] + zvals[:, :-1]) / 2
    diff = ((zvals[:, 1:] - zvals[:, :-1]) / 2) ** 2
    t1 = 3 * mid ** 2 + diff
    mu_t = mid + 2 * mid * diff / t1
    sig_t2 = diff / 3 - 4 * (diff ** 2) * (12 * mid ** 2 - diff) / 15 / (t1 ** 2)
    sig_r2 = (r ** 2) * (0.25 * mid ** 2 + 5 / 12 * diff - 4 * diff ** 2 / (15 * t1))
    d = rays[:, 3:]
    mu = rays[:, None, :3] + mu_t[:, :, None] * d[:, None, :]
    dd = d * d
    i_m = 1.0 - dd / d.norm()                                   # batch-global norm, :31
    diag = sig_t2[:, :, None] * dd[:, None, :] + sig_r2[:, :, None] * i_m[:, None, :]
    scales = torch.tensor([2.0 ** i for i in range(levels)], dtype=zvals.dtype, device=zvals.device)
    mu_r = scales[None, None, :, None] * mu[:, :, None, :]                       # (R,C,L,3)
    damp = torch.exp(-0.5 * (scales ** 2)[None, None, :, None] * diag[:, :, None, :])
    feat = torch.cat((torch.sin(mu_r) * damp, torch.cos(mu_r) * damp), dim=-1)   # (R,C,L,6)
    return feat.reshape(zvals.shape[0], zvals.shape[1] - 1, 6 * levels), mu, mu_t


# ----------------------------------------------------------------------------------------------
# f1  training-side callers                                   nerf/utils.py:72-94, nerf/addtional.py:14-18
# ----------------------------------------------------------------------------------------------
def valid_sampler(rgbs, coords, cam_tf, indices, jitter, point_num, focal, near, far):
    out_rgb = rgbs[indices]
    sc = coords[indices].to(torch.float32) + 0.5
    fx, fy = (focal[1], focal[0]) if isinstance(focal, (tuple, list)) else (focal, focal)
    sc = torch.stack((sc[..., 0] / fx, sc[..., 1] / fy), dim=-1)
    ray_raw = torch.sum(torch.cat([sc, -torch.ones(sc.shape[0], 1)], dim=-1).unsqueeze(-2) * cam_tf[:, :-1], dim=-1)
    res = (far - near) / point_num
    lengths = torch.linspace(near, far - res, point_num) + jitter * res
    pts = cam_tf[:, -1] + ray_raw[:, None, :] * lengths[:, :, None]
    rays = torch.cat((cam_tf[:, -1].unsqueeze(0).expand(ray_raw.shape[0], -1), ray_raw), dim=-1)
    return pts, lengths, out_rgb, rays


def get_bounds(weights, inds):
    starts, ends = inds[:, :-1], inds[:, 1:] + 1
    sat = torch.cat((torch.zeros(weights.shape[0], 1, device=weights.device), torch.cumsum(weights, dim=-1)), dim=-1)
    return torch.gather(sat, -1, ends) - torch.gather(sat, -1, starts)


# ----------------------------------------------------------------------------------------------
# f1  one training step of the non-Ref model: the run() closure + loss.backward()      train.py:164-199,206
# ----------------------------------------------------------------------------------------------
def train_step(sd_prop, sd_nerf, coarse_pts, coarse_lengths, rgb_targets, rays, u, n_fine=128, blur_alpha=0.01):
    """Losses and parameter gradients by torch autograd over the restated functions (which is what the reference does)."""
    sp = {k: v.detach().clone().requires_grad_(True) for k, v in sd_prop.items()}
    sn = {k: v.detach().clone().requires_grad_(True) for k, v in sd_nerf.items()}
    density = F.softplus(proposal_forward(sp, coarse_pts))                               # :166,169
    w_raw = weights_from_sigma(density, coarse_lengths, rays[:, 3:])                     # :170
    w_p = max_blur(w_raw, blur_alpha)                                                    # :171
    fine, below = inverse_sample(w_p.detach(), coarse_lengths, u, sort=True)             # :174
    fine = fine[..., :-1]                                                                # :189
    rgbo = nerf_forward(sn, length2pts(rays, fine))                                      # :190-191
    comp = composite(rgbo, fine, rays[:, 3:], white_bkg=False)                           # :192
    bounds = get_bounds(w_p, below)                                                      # :193
    img_loss = torch.mean((comp["rgb"] - rgb_targets) ** 2)                              # :195 (SoftL1Loss is MSE, addtional.py:38-43)
    wd = comp["weights"].detach()
    prop_loss = torch.sum(torch.relu(wd - bounds) ** 2 / (wd + 1e-8))                    # :197, addtional.py:20-24
    loss = prop_loss + img_loss                                                          # :198
    loss.backward()
    return {"loss": loss.detach(), "img_loss": img_loss.detach(), "prop_loss": prop_loss.detach(), "rendered": comp["rgb"].detach(),
            "bounds": bounds.detach(), "weights": wd, "fine": fine.detach(), "below": below,
            "grad_prop": {k: v.grad for k, v in sp.items()}, "grad_nerf": {k: v.grad for k, v in sn.items()}}


# ----------------------------------------------------------------------------------------------
# the per-ray path of render_image                             nerf/procedures.py:64-85
# ----------------------------------------------------------------------------------------------
def render_rays(sd_prop, sd_nerf, rays, base_z, jitter, u, near, far, n_fine=128, white_bkg=False, resolution=None,
                blur_alpha=0.01, softplus=False, pos_levels=10, dir_levels=4, torch_sum=False, chunk=4096):
    """rays (R,6), jitter (R,Pc), u (R,n_fine+1) -> dict with every intermediate the tests compare."""
    resolution = (far - near) / n_fine if resolution is None else resolution
    outs = {k: [] for k in ("rgb", "depth", "acc", "z_coarse", "sigma_prop", "sigma_prop_raw", "z_fine", "below", "weights", "rgbo")}
    for s in range(0, rays.shape[0], chunk):
        r, j, uu = rays[s:s + chunk], jitter[s:s + chunk], u[s:s + chunk]
        z_c = base_z + j * resolution                                             # :65
        pts = r[:, None, :3] + z_c[..., None] * r[:, None, 3:]                    # :66
        sigma_p = proposal_forward(sd_prop, pts, pos_levels)                      # :67
        sigma_raw = sigma_p
        if softplus:
            sigma_p = F.softplus(sigma_p)                                         # train.py:169
        w_p = max_blur(weights_from_sigma(sigma_p, z_c, r[:, 3:]), blur_alpha)    # :68-69
        z_f, below = inverse_sample(w_p, z_c, uu, sort=True, torch_sum=torch_sum)  # :70
        z_keep = z_f[..., :-1]                                                    # :76
        rgbo = nerf_forward(sd_nerf, length2pts(r, z_keep), pos_levels, dir_levels)  # :77-78
        comp = composite(rgbo, z_keep, r[:, 3:], white_bkg, (near, far))          # :80-85
        for k, v in (("rgb", comp["rgb"]), ("depth", comp["depth"]), ("acc", comp["acc"]), ("z_coarse", z_c),
                     ("sigma_prop", sigma_p), ("sigma_prop_raw", sigma_raw), ("z_fine", z_keep), ("below", below), ("weights", comp["weights"]),
                     ("rgbo", rgbo)):
            outs[k].append(v)
    return {k: torch.cat(v, dim=0) for k, v in outs.items()}


# ----------------------------------------------------------------------------------------------
# Philox4x32-10 restatement (the engine's own device RNG; Salmon et al., SC'11) for stream parity
# ----------------------------------------------------------------------------------------------
def philox_uniform(seed, ray_ids, n_samples, stream):
    """Uniforms the CUDA kernels draw for (seed, ray, sample, stream): returns (len(ray_ids), n_samples) fp32."""
    ray = np.asarray(ray_ids, dtype=np.uint64)[:, None]
    smp = np.arange(n_samples, dtype=np.uint64)[None, :]
    c = [np.broadcast_to(ray & np.uint64(0xFFFFFFFF), (ray.shape[0], n_samples)).copy(),
         np.broadcast_to(ray >> np.uint64(32), (ray.shape[0], n_samples)).copy(),
         np.broadcast_to(smp >> np.uint64(2), (ray.shape[0], n_samples)).copy(),
         np.full((ray.shape[0], n_samples), stream, dtype=np.uint64)]
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    M0, M1, W0, W1, m32 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k0, p1 & m32, (p0 >> np.uint64(32)) ^ c[3] ^ k1, p0 & m32]
        k0, k1 = (k0 + W0) & m32, (k1 + W1) & m32
    sel = np.broadcast_to(smp & np.uint64(3), c[0].shape)
    x = np.choose(sel.astype(np.int64), c)
    return torch.from_numpy(((x >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)))


# ---- Ref-NeRF directional front end (forward): integrated directional encoding, sRGB ---------------------
def ide_tables(deg_view):
    """(m, l) pairs and the (l_max + 1) x n_pairs coefficient matrix of /root/reference/nerf/ref_func.py:38-76
    (associated-Legendre / spherical-harmonic coefficients, float64 arithmetic rounded to fp32 on assignment)."""
    import math
    if deg_view > 5:
        raise ValueError("Only deg_view of at most 5 is numerically stable.")   # ref_func.py:67-68
    ml = [(m, 2 ** i) for i in range(deg_view) for m in range(2 ** i + 1)]     # ref_func.py:38-49
    l_max = 2 ** (deg_view - 1)

    def gen_binom(a, k):                                                       # ref_func.py:10-12
        return float(np.prod(a - np.arange(k))) / math.factorial(k)

    def legendre(l, m, k):                                                     # ref_func.py:14-31
        return ((-1) ** m * 2 ** l * math.factorial(l) / math.factorial(k) / math.factorial(l - k - m) *
                gen_binom(0.5 * (l + k + m - 1.0), l))

    def sph(l, m, k):                                                          # ref_func.py:33-36
        return math.sqrt((2.0 * l + 1.0) * math.factorial(l - m) / (4.0 * math.pi * math.factorial(l + m))) * legendre(l, m, k)

    mat = torch.zeros(l_max + 1, len(ml))
    for i, (m, l) in enumerate(ml):
        for k in range(l - m + 1):
            mat[k, i] = sph(l, m, k)
    return ml, mat


def ide(xyz, kappa_inv, deg_view=4):
    """integrated_dir_enc_fn, /root/reference/nerf/ref_func.py:78-108: xyz (..., 3), kappa_inv (..., 1) -> (..., 2 n_pairs)."""
    ml, mat = ide_tables(deg_view)
    mat = mat.to(device=xyz.device, dtype=xyz.dtype)     # (float64 inputs: tolerance analysis with the same fp32 tables)
    x, y, z = xyz[..., 0:1], xyz[..., 1:2], xyz[..., 2:3]
    m_arr = torch.tensor([m for m, _ in ml], device=xyz.device)
    l_arr = torch.tensor([l for _, l in ml], device=xyz.device)
    vmz = torch.cat([z ** i for i in range(mat.shape[0])], dim=-1)
    vmxy = torch.cat([(x + 1j * y) ** m for m in m_arr], dim=-1)
    sph_harms = vmxy * (vmz @ mat)
    sigma = 0.5 * l_arr * (l_arr + 1)
    out = sph_harms * torch.exp(-sigma * kappa_inv)
    return torch.cat([torch.real(out), torch.imag(out)], dim=-1)


def linear_to_srgb(linear):
    """/root/reference/nerf/nerf_helper.py:50-56."""
    eps = torch.full((1,), torch.finfo(torch.float32).eps, device=linear.device)
    srgb0 = 323 / 25 * linear
    srgb1 = (211 * torch.maximum(eps, linear) ** (5 / 12) - 11) / 200
    return torch.where(linear <= 0.0031308, srgb0, srgb1)


# ----------------------------------------------------------------------------------------------
# f3  Ref-NeRF forward (eval mode: no bottleneck noise)                        nerf/ref_model.py:67-109
# ----------------------------------------------------------------------------------------------
def refnerf_forward(sd, pts6, pos_levels=10, sh_level=4, use_srgb=False, relu_masks=None):
    """pts6 (..., 6) = [xyz, dir] -> (rgbo (..., 4) = [rgb, density], normal (..., 3)).
    relu_masks: see _relu (16 patterns: spa_block1, spa_block2, dir_block1, dir_block2, four layers each)."""
    counter = [0]

    def seq(h, prefix, idxs):
        for i in idxs:
            h = _relu(F.linear(h, sd[f"{prefix}.{i}.weight"], sd[f"{prefix}.{i}.bias"]), relu_masks, counter[0])
            counter[0] += 1
        return h
    x, ray_d = pts6[..., :3], pts6[..., 3:6]
    enc_x = torch.cat((x, positional_encoding(x, pos_levels)), dim=-1)                                  # :68-73
    x_tmp = seq(enc_x, "spa_block1", (0, 2, 4, 6))                                                      # :75
    inter = seq(torch.cat((enc_x, x_tmp), dim=-1), "spa_block2", (0, 2, 4, 6))                           # :76-77
    nct = F.linear(inter, sd["norm_col_tint_head.weight"], sd["norm_col_tint_head.bias"])               # :79
    normal, diffuse, tint = nct[..., :3], nct[..., 3:6], nct[..., 6:9]
    rt = F.linear(inter, sd["rho_tau_head.weight"], sd["rho_tau_head.bias"])                            # :80
    roughness, density = F.softplus(rt[..., :1] - 1.0), rt[..., 1:2]                                    # :81
    b = F.linear(inter, sd["bottle_neck.weight"], sd["bottle_neck.bias"])                               # :82
    normal = -normal / (normal.norm(dim=-1, keepdim=True) + 1e-7)                                       # :86
    reflect = ray_d - 2.0 * torch.sum(ray_d * normal, dim=-1, keepdim=True) * normal                    # :89
    wr_ide = ide(reflect, roughness, sh_level)                                                          # :90
    nv_dot = torch.sum(normal * ray_d, dim=-1, keepdim=True)                                            # :92
    all_in = torch.cat((b, wr_ide, nv_dot), dim=-1)                                                     # :94
    r_tmp = seq(all_in, "dir_block1", (0, 2, 4, 6))                                                     # :95
    q = seq(torch.cat((all_in, r_tmp), dim=-1), "dir_block2", (0, 2, 4, 6))                              # :96-98
    spec = torch.sigmoid(F.linear(q, sd["spec_rgb_head.0.weight"], sd["spec_rgb_head.0.bias"])) * torch.sigmoid(tint)
    if use_srgb:                                                                                         # :99-104
        rgb = linear_to_srgb(spec + torch.sigmoid(diffuse - math.log(3.0)))
    else:
        rgb = spec + torch.sigmoid(diffuse)
    return torch.cat((rgb, density), dim=-1), normal                                                    # :105
