"""GPU: Ref-NeRF forward (BASELINE configs[3]) against outputs of the UNMODIFIED reference
(tests/golden/make_golden.py round2: nerf/ref_model.py:67-109, the Ref branch of nerf/procedures.py:71-90,
nerf/nerf_base.py:58-73 with index bookkeeping) and against the oracle at larger sizes."""
import math

import pytest
import torch

import nerf_b200
from nerf_b200 import NeRF, ops
from oracle import nerf_oracle as O
from tests.golden.make_golden import inputs_ops, refnerf_inputs, render_case

pytestmark = pytest.mark.gpu
DEV = "cuda"


def refnet(use_srgb=False, precision=None):
    rn = nerf_b200.RefNeRF(10, 4, use_srgb=use_srgb)
    rn.load_state_dict(O.det_state_dict(rn, 7, gain=1.0))
    rn = rn.to(DEV).eval()
    rn.precision = precision
    return rn


@pytest.mark.parametrize("use_srgb", [False, True])
def test_refnerf_forward_vs_reference(golden_round2, use_srgb):
    rn = refnet(use_srgb)
    pts = refnerf_inputs()["pts"].to(DEV)
    with torch.no_grad():
        rgbo, normal = rn.forward(pts)
    g_rgbo = golden_round2["ref_fwd_rgbo_srgb" if use_srgb else "ref_fwd_rgbo"]
    assert rgbo.shape == g_rgbo.shape and normal.shape == golden_round2["ref_fwd_normal"].shape
    e_rgb = float((rgbo.cpu()[..., :3] - g_rgbo[..., :3]).abs().max())
    e_sig = float((rgbo.cpu()[..., 3] - g_rgbo[..., 3]).abs().max()) / max(1.0, float(g_rgbo[..., 3].abs().max()))
    e_n = float((normal.cpu() - golden_round2["ref_fwd_normal"]).abs().max())
    print("Ref-NeRF forward vs reference: rgb", e_rgb, "density (rel)", e_sig, "normal", e_n)
    # bf16 hi + lo operands carry 16 bits (2^-17 relative per product) through 17 layers
    assert e_rgb <= 5e-5 and e_sig <= 5e-5 and e_n <= 5e-5      # measured 1.4e-5 / 1.8e-5 / 1.7e-5


def test_refnerf_forward_vs_oracle_at_scale():
    rn = refnet()
    n = 512 * 192                      # BASELINE configs[3]: a 512-ray batch, 192 samples per ray
    pts = torch.cat((O.det_uniform((512, 192, 3), 71, -1.5, 1.5), O.det_uniform((512, 1, 3), 72, -1.0, 1.0).expand(512, 192, 3)), dim=-1).contiguous().to(DEV)
    with torch.no_grad():
        rgbo, normal = rn.forward(pts)
        sd = {k: v.to(DEV) for k, v in O.det_state_dict(nerf_b200.RefNeRF(10, 4), 7, gain=1.0).items()}
        r_rgbo, r_normal = O.refnerf_forward(sd, pts)
    assert float((rgbo[..., :3] - r_rgbo[..., :3]).abs().max()) <= 3e-4
    assert float((normal - r_normal).abs().max()) <= 3e-4
    for precision, tol in (("bf16", 8e-2),):
        rn.precision = precision
        with torch.no_grad():
            fast, _ = rn.forward(pts)
        assert float((fast[..., :3] - r_rgbo[..., :3]).abs().max()) <= tol
    assert n == rgbo.shape[0] * rgbo.shape[1]


def test_coarse_fine_merge_index_bookkeeping(golden_round2):
    go = inputs_ops()
    w = ops.max_blur(ops.weights_from_sigma(go["sigma"].to(DEV), go["z"].to(DEV), go["dirs"].to(DEV)), 0.01)
    zs, bs = ops.inverse_sample(w, go["z"].to(DEV), 129, sort=True, u=go["u"].to(DEV))
    pts, z, all_inds, sort_inds = NeRF.coarseFineMerge(go["rays"].to(DEV), go["z"].to(DEV), zs, bs)
    g = golden_round2
    assert z.shape == g["merge_z2"].shape and float((z.cpu() - g["merge_z2"]).abs().max()) < 1e-5
    # the permutation is the stable sort of cat(fine, coarse): identical unless two depths tie within an ulp
    same = (sort_inds.cpu() == g["merge_sort"]).float().mean()
    assert float(same) > 0.995, float(same)
    zz = torch.cat((zs, go["z"].to(DEV)), dim=-1)
    assert torch.equal(torch.gather(zz, -1, sort_inds), z)
    cat_inds = torch.cat((bs, torch.arange(64, device=DEV).expand(bs.shape[0], -1)), dim=-1)
    full_sort = torch.sort(zz, dim=-1, stable=True)[1]
    assert torch.equal(all_inds[:, :-1], torch.gather(cat_inds, -1, sort_inds))
    assert torch.equal(all_inds, torch.gather(cat_inds, -1, full_sort)) or float((all_inds == torch.gather(cat_inds, -1, full_sort)).float().mean()) > 0.995
    assert float((all_inds.cpu() == g["merge_inds"]).float().mean()) > 0.99


def test_render_image_ref_branch_vs_reference(golden_round2):
    """The Ref branch of render_image on the reference's 50x50 tile: rgb, depth and normal images."""
    H = W = 50
    pose, jitter, u, focal = render_case(H, W)
    rn = refnet()
    prop = nerf_b200.ProposalNetwork(10, 256)
    prop.load_state_dict(O.make_params("proposal", 1, "smooth"))
    prop = prop.to(DEV)
    prop.precision = "fp16x3"
    res = nerf_b200.render_image(rn, prop, pose.to(DEV), (H, W), focal, 2.0, 6.0, 128, white_bkg=True, render_depth=True, render_normal=True,
                                 jitter=jitter.to(DEV), u=u.to(DEV))
    g = golden_round2
    assert set(res) == {"rgb", "depth_img", "normal_img"} and res["rgb"].shape == (3, H, W)
    e_rgb = (res["rgb"].cpu() - g["ref_img_rgb"]).abs().amax(0)
    e_dep = (res["depth_img"][0].cpu() - g["ref_img_depth"]).abs()
    e_nrm = (res["normal_img"][0].cpu() - g["ref_img_normal"]).abs()
    mse = float(((res["rgb"].cpu() - g["ref_img_rgb"]) ** 2).mean())
    psnr = 99.0 if mse == 0 else -10.0 * math.log10(mse)
    print("Ref branch vs reference: rgb max", float(e_rgb.max()), "frac > 1e-3", float((e_rgb > 1e-3).float().mean()), "depth max", float(e_dep.max()),
          "normal max", float(e_nrm.max()), "PSNR", psnr)
    # measured: rgb max 1.4e-4, depth 8e-5, normal 4e-5, PSNR 102 dB (a handful of rays move a fine sample, tests/parity_tools.py)
    assert psnr > 90.0 and float(e_rgb.max()) <= 5e-4 and float(e_dep.max()) <= 3e-4 and float(e_nrm.max()) <= 2e-4
    assert float((e_rgb > 1e-4).float().mean()) <= 0.01
    # the MipNeRF path ignores render_normal exactly like the reference (procedures.py:41-42)
    net = nerf_b200.MipNeRF(10, 4, 256)
    net.load_state_dict(O.make_params("nerf", 2, "smooth"))
    out = nerf_b200.render_image(net.to(DEV), prop, pose.to(DEV), (H, W), focal, 2.0, 6.0, 128, render_normal=True, jitter=jitter.to(DEV), u=u.to(DEV))
    assert set(out) == {"rgb"}


def test_refnerf_launch_plan_reuse():
    """The recorded launch plan of a batch size is replayed on later calls: other inputs of the same size, a weight update and
    the training-mode perturbation path give what a freshly built module gives; outputs of earlier calls stay intact."""
    rn = refnet()
    a = O.det_uniform((64, 32, 6), 81, -1.0, 1.0).to(DEV)
    b = O.det_uniform((64, 32, 6), 82, -1.0, 1.0).to(DEV)
    with torch.no_grad():
        ya, na_ = rn.forward(a)
        keep = ya.clone()
        yb, nb_ = rn.forward(b)
        assert len(rn.__dict__["_nb2_ref_plans"]) == 1 and torch.equal(ya, keep)
        fresh = refnet()
        yb2, nb2_ = fresh.forward(b)
        assert torch.equal(yb, yb2) and torch.equal(nb_, nb2_)
        # a weight update is picked up (weights are re-converted in place, the plan's pointers stay valid)
        for m in (rn, fresh):
            m.spa_block1[0].weight.mul_(1.01)
            m.rho_tau_head.bias.add_(0.01)
        yc, _ = rn.forward(a)
        yc2, _ = refnet_like(fresh).forward(a)
        assert torch.equal(yc, yc2) and not torch.equal(yc, keep)
        # three batch sizes: the oldest plan is dropped, results unaffected
        for n in (16, 48):
            rn.forward(O.det_uniform((n, 32, 6), 83, -1.0, 1.0).to(DEV))
        assert len(rn.__dict__["_nb2_ref_plans"]) == rn.max_plans
        ya3, _ = rn.forward(a)
        assert torch.equal(ya3, yc)
        rn.train()
        torch.manual_seed(5)
        t1, _ = rn.forward(a)
        torch.manual_seed(5)
        t2, _ = rn.forward(a)
        assert torch.equal(t1, t2) and not torch.equal(t1, yc)


def refnet_like(src):
    """A new module holding src's current weights (no engine state)."""
    rn = nerf_b200.RefNeRF(10, 4)
    rn.load_state_dict(src.state_dict())
    return rn.to(DEV).eval()
