"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/nerf) on the CPU.

Run in the build container only (`python tests/golden/make_golden.py`); the GPU box has no
/root/reference, which is why the outputs are committed.  Import shims (no reference file is
edited): a `natsort` stub (nerf/dataset.py:11 imports it, not installed), `.cuda()` -> identity
(hard-coded .cuda() calls at nerf/nerf_base.py:81,85, nerf/addtional.py:103,106,
nerf/utils.py:118-129, nerf/procedures.py:50,65), and an injectable `torch.rand` so the CPU
draws at nerf/procedures.py:65 and nerf/utils.py:115 use known uniforms.

Inputs are NOT stored: they come from oracle.nerf_oracle.det_uniform / make_params (integer
hash -> exact fp32), which the tests call again with the same seeds.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.nerf_oracle import det_uniform, make_params  # noqa: E402  (input generators only)

REF = "/root/reference"


def import_reference():
    sys.modules.setdefault("natsort", types.SimpleNamespace(natsorted=sorted))
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    sys.path.insert(0, REF)
    import nerf.procedures as procedures
    import nerf.nerf_helper as nerf_helper
    import nerf.mip_methods as mip_methods
    import nerf.utils as utils
    from nerf.nerf_base import NeRF
    from nerf.mip_model import MipNeRF
    from nerf.addtional import ProposalNetwork
    return types.SimpleNamespace(procedures=procedures, nerf_helper=nerf_helper, mip_methods=mip_methods, utils=utils,
                                 NeRF=NeRF, MipNeRF=MipNeRF, ProposalNetwork=ProposalNetwork)


class RandQueue:
    """Stand-in for torch.rand that returns pre-drawn tensors in call order."""

    def __init__(self):
        self.q = []
        self.real = torch.rand

    def push(self, t):
        self.q.append(t)

    def __call__(self, *size, **kw):
        if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
            size = tuple(size[0])
        t = self.q.pop(0)
        assert tuple(t.shape) == tuple(size), (t.shape, size)
        return t.clone()


# ---- shared input definitions (tests/golden_inputs.py re-uses these) ---------------------------
def inputs_ops():
    R, P, N = 24, 64, 129
    g = {}
    g["pe_x"] = det_uniform((40, 3), 11, -6.0, 6.0)
    g["pe_x3"] = det_uniform((5, 7, 3), 12, -1.0, 1.0)
    z = torch.linspace(2.0, 6.0, P)[None, :] + det_uniform((R, P), 13, 0.0, 1.0) * (4.0 / 128)
    g["z"] = z
    sig = det_uniform((R, P), 14, -20.0, 40.0)
    sig[0] = 0.0            # empty ray
    sig[1] = -5.0           # all-negative density
    sig[2] = 500.0          # opaque at the first sample
    sig[3, :40] = -1.0      # single spike -> exercises the denom < 1e-5 branch of sample_pdf
    sig[3, 41:] = -1.0
    sig[3, 40] = 2000.0
    g["sigma"] = sig
    g["dirs"] = det_uniform((R, 3), 15, -1.0, 1.0)
    u = det_uniform((R, N), 16, 0.0, 1.0)
    u[4, 0] = 0.0
    u[4, 1] = 1.0 - 2.0 ** -24
    g["u"] = u
    g["rays"] = torch.cat((det_uniform((R, 3), 17, -1.0, 1.0) * 0.5 + torch.tensor([0.0, 0.0, 4.0]), g["dirs"]), dim=-1)
    g["rgbo"] = torch.cat((det_uniform((R, 128, 3), 18, 0.0, 1.0), det_uniform((R, 128, 1), 19, -10.0, 60.0)), dim=-1)
    zf = torch.sort(det_uniform((R, 128), 20, 2.0, 6.0), dim=-1)[0]
    g["z_fine"] = zf
    g["ipe_z"] = torch.sort(det_uniform((8, 17), 21, 2.0, 6.0), dim=-1)[0]
    g["ipe_rays"] = g["rays"][:8].clone()
    g["mlp_pts"] = torch.cat((det_uniform((6, 16, 3), 22, -1.5, 1.5), det_uniform((6, 1, 3), 23, -1.0, 1.0).expand(6, 16, 3)), dim=-1).contiguous()
    return g


def inputs_train():
    """One 40x30 'training image' worth of pixels for validSampler."""
    Hh, Ww = 30, 40
    from nerf_b200.utils import pose_spherical
    rows, cols = torch.meshgrid(torch.arange(Hh), torch.arange(Ww), indexing="ij")
    coords = torch.stack((cols - Ww // 2, Hh // 2 - rows), dim=-1).reshape(-1, 2)
    return {"rgbs": det_uniform((Hh * Ww, 3), 61, 0.0, 1.0), "coords": coords, "cam_tf": pose_spherical(-40.0, -30.0, 4.0)[:3, :].contiguous(),
            "indices": (torch.arange(96) * 37 + 5) % (Hh * Ww), "jitter": det_uniform((96, 64), 62, 0.0, 1.0), "focal": (55.0, 60.0)}


def render_case(H, W):
    from nerf_b200.utils import pose_spherical  # same formula as the reference; pure host math
    pose = pose_spherical(30.0, -30.0, 4.0)[:3, :].contiguous()
    jitter = det_uniform((H * W, 64), 31, 0.0, 1.0)
    u = det_uniform((H * W, 129), 32, 0.0, 1.0)
    focal = float(W / np.tan(0.5 * 0.6911112070083618))     # fov2Focal with a scalar fov (utils.py:102-105)
    return pose, jitter, u, focal


def load_sd(module, sd):
    module.load_state_dict({k: v.clone() for k, v in sd.items()})
    return module


def main():
    ref = import_reference()
    rq = RandQueue()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    g = inputs_ops()
    out = {}
    with torch.no_grad():
        out["pe"] = ref.nerf_helper.positional_encoding(g["pe_x"], 10)
        out["pe3"] = ref.nerf_helper.positional_encoding(g["pe_x3"], 4)
        w_raw = ref.ProposalNetwork.get_weights(g["sigma"], g["z"], g["dirs"])
        out["weights"] = w_raw
        out["weights_nodir"] = ref.NeRF.getNormedWeight(g["sigma"], g["z"])
        w_blur = ref.mip_methods.maxBlurFilter(w_raw, 0.01)
        out["blur"] = w_blur
        mids = 0.5 * (g["z"][:, 1:] + g["z"][:, :-1])
        torch.rand = rq
        rq.push(g["u"])
        s, b, a = ref.utils.sample_pdf(mids, w_blur[:, 1:-1], 129)
        out["pdf_samples"], out["pdf_below"], out["pdf_above"] = s, b, a
        rq.push(g["u"])
        zs, bs = ref.utils.inverseSample(w_blur, g["z"], 129, sort=True)
        out["inv_z"], out["inv_below"] = zs, bs
        rq.push(g["u"])
        out["inv_z_unsorted"] = ref.utils.inverseSample(w_blur, g["z"], 129, sort=False)
        torch.rand = rq.real
        # the cdf the reference builds internally (utils.py:110-113), for the search-only stage
        wp = w_blur[:, 1:-1] + 1e-5
        pdf = wp / torch.sum(wp, -1, keepdim=True)
        cdf = torch.cumsum(pdf, -1)
        out["cdf"] = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
        out["l2p"] = ref.NeRF.length2pts(g["rays"], g["z_fine"])
        mp, mz = ref.NeRF.coarseFineMerge(g["rays"], g["z"], zs)
        out["merge_pts"], out["merge_z"] = mp, mz
        rgb, w, ex = ref.NeRF.render(g["rgbo"], g["z_fine"], g["dirs"], white_bkg=True, render_depth=(2.0, 6.0))
        out["comp_rgb"], out["comp_w"], out["comp_depth"] = rgb, w, ex["depth_img"]
        rgb2, _, _ = ref.NeRF.render(g["rgbo"], g["z_fine"], g["dirs"], white_bkg=False)
        out["comp_rgb_black"] = rgb2
        f, mu, mu_t = ref.mip_methods.ipe_feature(g["ipe_z"], g["ipe_rays"], 10, 0.01)
        out["ipe_feat"], out["ipe_mu"], out["ipe_mu_t"] = f, mu, mu_t

        # training-side callers: validSampler (utils.py:72-94) with injected randint / rand, getBounds (addtional.py:14-18)
        import nerf.addtional as addtional
        vs = inputs_train()
        real_randint = torch.randint
        torch.randint = lambda *a, **k: vs["indices"].clone()
        torch.rand = rq
        rq.push(vs["jitter"])
        vp, vl, vrgb, vrays = ref.utils.validSampler(vs["rgbs"], vs["coords"], vs["cam_tf"], vs["indices"].numel(), 64, vs["focal"], 2.0, 6.0, True)
        torch.rand = rq.real
        torch.randint = real_randint
        out["vs_pts"], out["vs_len"], out["vs_rgb"], out["vs_rays"] = vp, vl, vrgb, vrays
        out["bounds"] = addtional.getBounds(w_blur, bs)

        # integrated positional encoding fed through the proposal network's encoded_pt hook (addtional.py:88-91)
        prop_ipe = load_sd(ref.ProposalNetwork(10, 256), make_params("proposal", 1, "smooth"))
        out["prop_fwd_ipe"] = prop_ipe.forward(mu, encoded_pt=f)

        for style in ("he", "smooth", "refinit"):
            prop = load_sd(ref.ProposalNetwork(10, 256), make_params("proposal", 1, style))
            net = load_sd(ref.MipNeRF(10, 4, 256), make_params("nerf", 2, style))
            out[f"prop_fwd_{style}"] = prop.forward(g["mlp_pts"][..., :3].contiguous())
            out[f"nerf_fwd_{style}"] = net.forward(g["mlp_pts"])
            # render_image on one 50x50 tile (the reference's own tile size)
            H = W = 50
            pose, jitter, u, focal = render_case(H, W)
            torch.rand = rq
            rq.push(jitter.view(H, W, 64))
            rq.push(u)
            res = ref.procedures.render_image(net, prop, pose, (H, W), focal, 2.0, 6.0, 128, white_bkg=True, render_depth=True)
            torch.rand = rq.real
            out[f"img_rgb_{style}"] = res["rgb"]
            out[f"img_depth_{style}"] = res["depth_img"][0]

    # config 1 (64x64, 32 coarse): the reference's render_image crashes there (procedures.py:24-31), so the
    # golden is the trainer's composition (train.py:160-192) on 256 rays of a 64x64 image, eval-style.
    with torch.no_grad():
        H = W = 64
        pose, _, _, focal = render_case(H, W)
        focal = float(W / np.tan(0.5 * 0.6911112070083618))
        prop = load_sd(ref.ProposalNetwork(10, 256), make_params("proposal", 1, "he"))
        net = load_sd(ref.MipNeRF(10, 4, 256), make_params("nerf", 2, "he"))
        Rn, Pc = 256, 32
        pix = (torch.arange(Rn) * 16) % (H * W)
        rows, cols = pix // W, pix % W
        coords = torch.stack((cols - W // 2, H // 2 - rows), dim=-1).to(torch.float32) + 0.5
        coords = coords / focal
        ray_raw = torch.sum(torch.cat([coords, -torch.ones(Rn, 1)], dim=-1).unsqueeze(-2) * pose[:, :-1], dim=-1)
        rays = torch.cat((pose[:, -1].unsqueeze(0).expand(Rn, -1), ray_raw), dim=-1)
        resolution = (6.0 - 2.0) / Pc
        lengths = torch.linspace(2.0, 6.0 - resolution, Pc) + det_uniform((Rn, Pc), 41, 0.0, 1.0) * resolution  # utils.py:87-89
        pts = pose[:, -1] + ray_raw[:, None, :] * lengths[:, :, None]
        density = torch.nn.functional.softplus(prop.forward(pts))                       # train.py:166,169
        w_blur = ref.mip_methods.maxBlurFilter(ref.ProposalNetwork.get_weights(density, lengths, rays[:, 3:]), 0.01)
        torch.rand = rq
        rq.push(det_uniform((Rn, 129), 42, 0.0, 1.0))
        fine, below = ref.utils.inverseSample(w_blur, lengths, 129, sort=True)
        torch.rand = rq.real
        fine = fine[..., :-1]
        rgbo = net.forward(ref.NeRF.length2pts(rays, fine))
        rgb, wts, _ = ref.NeRF.render(rgbo, fine, rays[:, 3:])
        out["c1_rays"], out["c1_lengths"] = rays, lengths
        out["c1_rgb"], out["c1_fine"], out["c1_below"], out["c1_density"] = rgb, fine, below, density

    arrays = {}
    for k, v in out.items():
        v = v.detach().cpu()
        arrays[k] = v.numpy().astype(np.int16) if v.dtype == torch.int64 else v.numpy()
    path = os.path.join(HERE, "reference_outputs.npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", len(arrays), "arrays; torch", torch.__version__)


# ---- Ref-NeRF directional front end: a second, separate file so that the pinned one is never rewritten ----
def inputs_refnerf():
    g = {}
    d = det_uniform((16, 16, 3), 41, -1.0, 1.0)
    g["ide_dirs"] = d / d.norm(dim=-1, keepdim=True)
    g["ide_kappa_inv"] = det_uniform((16, 16, 1), 42, 0.05, 1.0)          # roughness = softplus(. - 1) > 0 (ref_model.py:83)
    g["srgb_lin"] = det_uniform((512, 3), 43, -0.01, 1.2)
    return g


def main_refnerf():
    """nerf/ref_func.py needs `np.math` (removed in NumPy 2): shimmed with the stdlib module, no reference file is edited."""
    import math
    np.math = math
    torch.Tensor.cuda = lambda self, *a, **k: self
    sys.path.insert(0, REF)
    from nerf.ref_func import generate_ide_fn
    from nerf.nerf_helper import linear_to_srgb
    g = inputs_refnerf()
    out = {}
    for deg in (1, 4, 5):
        out[f"ide_deg{deg}"] = generate_ide_fn(deg)(g["ide_dirs"], g["ide_kappa_inv"])
    out["srgb"] = linear_to_srgb(g["srgb_lin"])
    path = os.path.join(HERE, "reference_outputs_refnerf.npz")
    np.savez_compressed(path, **{k: v.detach().cpu().numpy() for k, v in out.items()})
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", len(out), "arrays; torch", torch.__version__)


# ---- round 2: multi-tile render_image under torch.manual_seed, Ref-NeRF forward, one training step ----------------------
def refnerf_inputs():
    g = {}
    g["pts"] = torch.cat((det_uniform((6, 24, 3), 71, -1.5, 1.5), det_uniform((6, 1, 3), 72, -1.0, 1.0).expand(6, 24, 3)), dim=-1).contiguous()
    return g


def train_inputs():
    """One training batch (train.py:157-163): 48 rays of the 40x30 image of inputs_train(), 64 coarse + 128 fine."""
    vs = inputs_train()
    return {"vs": vs, "indices": (torch.arange(48) * 23 + 3) % (30 * 40), "jitter": det_uniform((48, 64), 81, 0.0, 1.0),
            "u": det_uniform((48, 129), 82, 0.0, 1.0)}


GRAD_HEAD = 48   # leading entries of every gradient tensor that are stored next to its norm and sum


def main_round2():
    import math
    np.math = math
    ref = import_reference()
    from nerf_b200.synthetic import det_state_dict
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    out = {}
    rq = RandQueue()
    with torch.no_grad():
        # (1) render_image over several tiles with the reference's OWN torch.rand draws (no injection): 100x100 = 2x2 tiles
        # of 50, and 70x100 where the reference leaves rows 50..69 unrendered (procedures.py:24-31,60-64)
        prop = load_sd(ref.ProposalNetwork(10, 256), make_params("proposal", 1, "smooth"))
        net = load_sd(ref.MipNeRF(10, 4, 256), make_params("nerf", 2, "smooth"))
        for (H, W) in ((100, 100), (70, 100)):
            pose, _, _, _ = render_case(H, W)
            focal = float(W / np.tan(0.5 * 0.6911112070083618))
            torch.manual_seed(2024)
            res = ref.procedures.render_image(net, prop, pose, (H, W), focal, 2.0, 6.0, 128, white_bkg=True, render_depth=True)
            out[f"seeded_rgb_{H}x{W}"] = res["rgb"]
            out[f"seeded_depth_{H}x{W}"] = res["depth_img"][0]

        # (2) Ref-NeRF (nerf/ref_model.py:16-118), eval mode (no bottleneck noise), deterministic parameters
        from nerf.ref_model import RefNeRF
        rn = RefNeRF(10, 4)
        rn.load_state_dict(det_state_dict(rn, 7, gain=1.0))
        rn.eval()
        g = refnerf_inputs()
        rgbo, normal = rn.forward(g["pts"])
        out["ref_fwd_rgbo"], out["ref_fwd_normal"] = rgbo, normal
        rn_srgb = RefNeRF(10, 4, use_srgb=True)
        rn_srgb.load_state_dict(det_state_dict(rn_srgb, 7, gain=1.0))
        rn_srgb.eval()
        out["ref_fwd_rgbo_srgb"] = rn_srgb.forward(g["pts"])[0]
        # coarseFineMerge with index bookkeeping (nerf_base.py:58-73)
        go = inputs_ops()
        w_blur = ref.mip_methods.maxBlurFilter(ref.ProposalNetwork.get_weights(go["sigma"], go["z"], go["dirs"]), 0.01)
        torch.rand = rq
        rq.push(go["u"])
        zs, bs = ref.utils.inverseSample(w_blur, go["z"], 129, sort=True)
        torch.rand = rq.real
        mp, mz, minds, msort = ref.NeRF.coarseFineMerge(go["rays"], go["z"], zs, bs)
        out["merge_inds"], out["merge_sort"], out["merge_z2"] = minds, msort, mz
        # the Ref branch of render_image on one 50x50 tile (procedures.py:71-74,80-90), normals and depth on
        H = W = 50
        pose, jitter, u, focal = render_case(H, W)
        torch.rand = rq
        rq.push(jitter.view(H, W, 64))
        rq.push(u)
        res = ref.procedures.render_image(rn, prop, pose, (H, W), focal, 2.0, 6.0, 128, white_bkg=True, render_depth=True, render_normal=True)
        torch.rand = rq.real
        out["ref_img_rgb"], out["ref_img_depth"], out["ref_img_normal"] = res["rgb"], res["depth_img"][0], res["normal_img"][0]

    # (3) one training step of the non-Ref model: the run() closure of train.py:164-199 executed with the reference's own
    # functions on injected draws, loss.backward(), gradient summaries of every parameter
    import nerf.addtional as addtional
    ti = train_inputs()
    vs = ti["vs"]
    prop = load_sd(ref.ProposalNetwork(10, 256), make_params("proposal", 1, "smooth"))
    net = load_sd(ref.MipNeRF(10, 4, 256), make_params("nerf", 2, "smooth"))
    real_randint = torch.randint
    torch.randint = lambda *a, **k: ti["indices"].clone()
    torch.rand = rq
    rq.push(ti["jitter"])
    coarse_samples, coarse_lengths, rgb_targets, coarse_cam_rays = ref.utils.validSampler(
        vs["rgbs"], vs["coords"], vs["cam_tf"], ti["indices"].numel(), 64, vs["focal"], 2.0, 6.0, True)
    torch.randint = real_randint
    density = torch.nn.functional.softplus(prop.forward(coarse_samples))                                   # train.py:166,169
    prop_weights_raw = ref.ProposalNetwork.get_weights(density, coarse_lengths, coarse_cam_rays[:, 3:])
    prop_weights = ref.mip_methods.maxBlurFilter(prop_weights_raw, 0.01)
    rq.push(ti["u"])
    fine_lengths, below_idxs = ref.utils.inverseSample(prop_weights, coarse_lengths, 129, sort=True)
    torch.rand = rq.real
    fine_lengths = fine_lengths[..., :-1]
    fine_samples = ref.NeRF.length2pts(coarse_cam_rays, fine_lengths)
    fine_rgbo = net.forward(fine_samples)
    fine_rendered, weights, _ = ref.NeRF.render(fine_rgbo, fine_lengths, coarse_cam_rays[:, 3:])
    weight_bounds = addtional.getBounds(prop_weights, below_idxs)
    img_loss = addtional.SoftL1Loss()(fine_rendered, rgb_targets)
    prop_loss = addtional.ProposalLoss()(weight_bounds, weights.detach())
    loss = prop_loss + img_loss
    loss.backward()
    out["train_loss"] = torch.stack((loss.detach(), img_loss.detach(), prop_loss.detach()))
    out["train_rendered"], out["train_bounds"], out["train_weights"] = fine_rendered.detach(), weight_bounds.detach(), weights.detach()
    for tag, m in (("prop", prop), ("nerf", net)):
        for k, p_ in m.named_parameters():
            gk = p_.grad.reshape(-1)
            out[f"grad_{tag}_{k}"] = torch.cat((gk.norm().reshape(1), gk.sum().reshape(1), gk[:GRAD_HEAD]))

    arrays = {}
    for k, v in out.items():
        v = v.detach().cpu()
        arrays[k] = v.numpy().astype(np.int16) if v.dtype == torch.int64 else v.numpy()
    path = os.path.join(HERE, "reference_outputs_round2.npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", len(arrays), "arrays; torch", torch.__version__)


# ---- round 2, late: the training side of Ref-NeRF (train.py:176-187 with is_ref_model) ----------------------------------
def ref_train_inputs():
    """6 rays x 24 merged samples: positions / directions of refnerf_inputs(), sorted sample depths, colour targets and a
    sort permutation whose last 8 pre-sort positions play the coarse samples (coarse_grad_select)."""
    g = refnerf_inputs()
    z = torch.sort(det_uniform((6, 24), 73, 2.0, 6.0), dim=-1).values
    sort_inds = torch.argsort(det_uniform((6, 24), 74, 0.0, 1.0), dim=-1)
    return {"pos": g["pts"][..., :3].contiguous(), "dirs": g["pts"][..., 3:].contiguous(), "z": z, "sort_inds": sort_inds,
            "targets": det_uniform((6, 3), 75, 0.0, 1.0)}


def main_round3():
    import math
    np.math = math
    ref = import_reference()
    from nerf_b200.synthetic import det_state_dict
    import nerf.addtional as addtional
    from nerf.ref_model import BackFaceLoss, RefNeRF, WeightedNormalLoss
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    out = {}
    ti = ref_train_inputs()
    rn = RefNeRF(10, 4)
    rn.load_state_dict(det_state_dict(rn, 7, gain=1.0))
    rn.eval()                                                    # (train mode adds torch.normal noise to the bottleneck, ref_model.py:86-87)
    fine_pos, fine_dir = ti["pos"].clone(), ti["dirs"]
    fine_pos.requires_grad = True                                                                          # train.py:177
    fine_rgbo, pred_normal = rn.forward(fine_pos, fine_dir)
    density_grad = -RefNeRF.get_grad(fine_rgbo[..., -1], fine_pos)                                         # train.py:180
    fine_rgbo[..., -1] = torch.nn.functional.softplus(fine_rgbo[..., -1] + 0.5)
    fine_rendered, weights, _ = ref.NeRF.render(fine_rgbo, ti["z"], fine_dir[:, 0], rn.density_act)        # train.py:182 (sic: 4th positional)
    normal_loss = WeightedNormalLoss(True)(weights, density_grad, pred_normal)
    bf_loss = BackFaceLoss()(weights, pred_normal, fine_dir)
    img_loss = addtional.SoftL1Loss()(fine_rendered, ti["targets"])
    loss = img_loss + 4e-4 * normal_loss + 0.1 * bf_loss
    loss.backward()
    out["rt_density_grad"], out["rt_weights"], out["rt_rendered"] = density_grad.detach(), weights.detach(), fine_rendered.detach()
    out["rt_losses"] = torch.stack((loss.detach(), img_loss.detach(), normal_loss.detach(), bf_loss.detach()))
    out["rt_pos_grad"] = fine_pos.grad.detach()
    out["rt_select"] = RefNeRF.coarse_grad_select(density_grad.detach(), ti["sort_inds"], 8)
    for k, p_ in rn.named_parameters():
        gk = p_.grad.reshape(-1)
        out[f"rt_grad_{k}"] = torch.cat((gk.norm().reshape(1), gk.sum().reshape(1), gk[:GRAD_HEAD]))
    arrays = {k: v.detach().cpu().numpy() for k, v in out.items()}
    path = os.path.join(HERE, "reference_outputs_ref_train.npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", len(arrays), "arrays; torch", torch.__version__)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "round3":
        main_round3()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "refnerf":
        main_refnerf()
    elif len(sys.argv) > 1 and sys.argv[1] == "round2":
        main_round2()
    else:
        main()
