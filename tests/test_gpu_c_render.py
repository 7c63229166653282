"""GPU: the fused render path (nb2_render_rays / render_image) against the reference's render_image
(golden 50x50 tile) and against the oracle at BASELINE.json's full sizes."""
import math

import pytest
import torch

import nerf_b200
from nerf_b200 import ops
from oracle import nerf_oracle as O
from tests.golden.make_golden import render_case

pytestmark = pytest.mark.gpu
DEV = "cuda"
FOV = 0.6911112070083618


def load(module, sd):
    module.load_state_dict({k: v.clone() for k, v in sd.items()})
    return module.to(DEV)


def nets(style="he", precision=None):
    prop = load(nerf_b200.ProposalNetwork(10, 256), O.make_params("proposal", 1, style))
    net = load(nerf_b200.MipNeRF(10, 4, 256), O.make_params("nerf", 2, style))
    prop.precision = net.precision = precision
    return net, prop


def psnr(a, b):
    mse = float(((a - b) ** 2).mean())
    return 99.0 if mse == 0 else -10.0 * math.log10(mse)


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
@pytest.mark.parametrize("style", ["smooth", "refinit"])
def test_render_image_matches_reference_tile(golden, precision, style):
    """north_star parity against the reference's own render_image output on identical rays and uniforms.

    refinit (the reference's fresh-model regime): RGB and depth within 1e-4 abs, every mode.
    smooth (band-limited random field, O(1) gains): the fp32 mode stays within 1e-4 on RGB (2e-4 on depth, whose
    value sums 128 products of magnitude ~5); the tensor-core fp32-faithful mode is bounded by the TMEM
    accumulator's truncating adds (relative ~1e-5 on the density), which the resampling step turns into depth
    shifts of ~1e-5 on a few rays: >= 98 % of pixels within 1e-4, none beyond 3e-3, PSNR vs the reference > 85 dB.
    """
    H = W = 50
    pose, jitter, u, focal = render_case(H, W)
    net, prop = nets(style, precision)
    res = nerf_b200.render_image(net, prop, pose.to(DEV), (H, W), focal, 2.0, 6.0, 128, white_bkg=True, render_depth=True,
                                 jitter=jitter.to(DEV), u=u.to(DEV))
    assert res["rgb"].shape == (3, H, W) and res["depth_img"].shape == (3, H, W)
    e_rgb = (res["rgb"].cpu() - golden[f"img_rgb_{style}"]).abs()
    e_dep = (res["depth_img"][0].cpu() - golden[f"img_depth_{style}"]).abs()
    frac = float((e_rgb.amax(dim=0) > 1e-4).float().mean())
    p = psnr(res["rgb"].cpu(), golden[f"img_rgb_{style}"])
    print(precision, style, "max rgb err", float(e_rgb.max()), "max depth err", float(e_dep.max()), "frac px > 1e-4", frac, "PSNR", p)
    if style == "refinit":
        assert float(e_rgb.max()) <= 1e-4 and float(e_dep.max()) <= 1e-4
    elif precision == "fp32":
        assert float(e_rgb.max()) <= 1e-4 and float(e_dep.max()) <= 2e-4
    else:
        assert frac <= 0.02 and float(e_rgb.max()) <= 3e-3 and p > 85.0


@pytest.mark.parametrize("precision", ["bf16", "fp16", "bf16x3"])
@pytest.mark.parametrize("style", ["smooth", "refinit"])
def test_render_image_reduced_precision_psnr(golden, style, precision):
    """The single-pass / bf16-split modes are judged by PSNR against the reference image, not by 1e-4."""
    H = W = 50
    pose, jitter, u, focal = render_case(H, W)
    net, prop = nets(style, precision)
    res = nerf_b200.render_image(net, prop, pose.to(DEV), (H, W), focal, 2.0, 6.0, 128, white_bkg=True,
                                 jitter=jitter.to(DEV), u=u.to(DEV))
    p = psnr(res["rgb"].cpu(), golden[f"img_rgb_{style}"])
    print(precision, style, "PSNR vs reference image", p)
    assert p > ({"bf16": 30.0, "fp16": 45.0, "bf16x3": 75.0}[precision] if style == "smooth" else 100.0)


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_chaotic_field_stagewise(golden, precision):
    """'he' field: a one-ulp depth change moves the reference's own RGB by > 1e-4 (test_oracle_golden.py), so
    parity is staged: proposal density close, and GIVEN the reference's fine depths the fine stage (encode +
    8x256 MLP + compositing) is within 1e-4."""
    H = W = 50
    pose, jitter, u, focal = render_case(H, W)
    sp, sn = O.make_params("proposal", 1, "he"), O.make_params("nerf", 2, "he")
    rays = O.generate_rays(pose, H, W, focal)
    ref = O.render_rays(sp, sn, rays, torch.linspace(2.0, 6.0, 64), jitter, u, 2.0, 6.0, 128, white_bkg=True)
    net, prop = nets("he", precision)
    with torch.no_grad():
        rgbo = net.forward(nerf_b200.NeRF.length2pts(rays.to(DEV), ref["z_fine"].to(DEV)))
        rgb, w, ex = nerf_b200.NeRF.render(rgbo, ref["z_fine"].to(DEV), rays[:, 3:].contiguous().to(DEV), white_bkg=True, render_depth=(2.0, 6.0))
    e = float((rgb.cpu() - ref["rgb"]).abs().max())
    ed = float((ex["depth_img"].cpu() - ref["depth"]).abs().max())
    print(precision, "fine stage on reference depths: max rgb err", e, "max depth err", ed)
    assert e <= 1e-4 and ed <= 1e-4
    # and the image the reference itself produced for this tile
    img = rgb.view(H, W, 3).permute(2, 0, 1).cpu()
    assert float((img - golden["img_rgb_he"]).abs().max()) <= 2e-3


def test_intermediates_and_sample_indices(golden):
    """Staged exactness: coarse depths bit-exact; fine depths agree with the oracle to 1e-5 except where a
    sample sits on a cdf knot (index flip), which must be rare."""
    H = W = 50
    pose, jitter, u, focal = render_case(H, W)
    net, prop = nets("smooth", "fp32")
    net._nb2_sync(); prop._nb2_sync()
    rays = ops.generate_rays(pose.to(DEV), H, W, focal, focal)
    base = torch.linspace(2.0, 6.0, 64, device=DEV)
    out = ops.render_rays(rays, base, 2.0, 6.0, 128, white_bkg=True, precision="fp32", jitter=jitter.to(DEV), u=u.to(DEV), debug=True)
    ref = O.render_rays(O.make_params("proposal", 1, "smooth"), O.make_params("nerf", 2, "smooth"), O.generate_rays(pose, H, W, focal),
                        base.cpu(), jitter, u, 2.0, 6.0, 128, white_bkg=True)
    assert torch.equal(out["z_coarse"].cpu(), ref["z_coarse"])
    assert float((out["sigma_prop"].cpu() - ref["sigma_prop"]).abs().max()) <= 1e-5 * max(50.0, float(ref["sigma_prop"].abs().max()))
    dz = (out["z_fine"].cpu() - ref["z_fine"]).abs()
    print("fine depths: median |dz|", float(dz.median()), "frac > 1e-4", float((dz > 1e-4).float().mean()))
    assert float(dz.median()) < 2e-6 and float((dz > 1e-4).float().mean()) < 2e-3
    zf = out["z_fine"]
    assert bool((zf[:, 1:] >= zf[:, :-1]).all())


@pytest.mark.parametrize("precision", ["fp16x3", "bf16", "fp16"])
def test_full_size_400x400_vs_oracle_on_device(precision):
    """Config 2 at full size: engine vs the oracle (PyTorch fp32 on the same GPU) on identical rays and uniforms."""
    H = W = 400
    pose = nerf_b200.pose_spherical(30.0, -30.0, 4.0)[:3, :].to(DEV)
    focal = nerf_b200.fov2Focal(FOV, (H, W))[0]
    g = torch.Generator(device="cpu").manual_seed(1234)
    jitter = torch.rand(H * W, 64, generator=g).to(DEV)
    u = torch.rand(H * W, 129, generator=g).to(DEV)
    net, prop = nets("smooth", precision)
    res = nerf_b200.render_image(net, prop, pose, (H, W), focal, 2.0, 6.0, 128, white_bkg=True, render_depth=True, jitter=jitter, u=u)
    sp, sn = O.params_to(O.make_params("proposal", 1, "smooth"), DEV), O.params_to(O.make_params("nerf", 2, "smooth"), DEV)
    rays = O.generate_rays(pose, H, W, focal)
    ref = O.render_rays(sp, sn, rays, torch.linspace(2.0, 6.0, 64, device=DEV), jitter, u, 2.0, 6.0, 128, white_bkg=True, chunk=8000)
    rgb = res["rgb"].permute(1, 2, 0).reshape(-1, 3)
    err = (rgb - ref["rgb"]).abs().max(dim=-1)[0]
    derr = (res["depth_img"][0].reshape(-1) - ref["depth"]).abs()
    p = psnr(rgb, ref["rgb"])
    frac = float((err > 1e-4).float().mean())
    print(precision, "400x400: PSNR vs oracle", p, "max rgb err", float(err.max()), "frac rays > 1e-4", frac, "max depth err", float(derr.max()))
    if precision == "fp16x3":
        # fp32-faithful mode: within 1e-4 except rays where a fine sample crosses a cdf knot / the
        # denom<1e-5 branch of sample_pdf (a discontinuity of the reference algorithm itself)
        assert frac < 2e-2 and p > 85.0
    else:
        assert p > {"bf16": 30.0, "fp16": 42.0}[precision]


def test_shard_invariance_and_ray_permutation():
    """Philox is keyed on the global ray index: rendering in pieces (any sharding) is bit-identical."""
    H = W = 120
    pose = nerf_b200.pose_spherical(-60.0, -30.0, 4.0)[:3, :].to(DEV)
    focal = nerf_b200.fov2Focal(FOV, (H, W))[0]
    net, prop = nets("he", "bf16")
    net._nb2_sync(); prop._nb2_sync()
    rays = ops.generate_rays(pose, H, W, focal, focal)
    base = torch.linspace(2.0, 6.0, 64, device=DEV)
    whole = ops.render_rays(rays, base, 2.0, 6.0, 128, white_bkg=True, precision="bf16", seed=42)
    parts = []
    for s, c in ((0, 4096), (4096, 5000), (9096, H * W - 9096)):
        parts.append(ops.render_rays(rays[s:s + c].contiguous(), base, 2.0, 6.0, 128, white_bkg=True, precision="bf16", seed=42, ray_offset=s)["rgb"])
    assert torch.equal(torch.cat(parts), whole["rgb"])
    assert float(whole["acc"].min()) >= 0.0 and float(whole["acc"].max()) <= 1.0 + 1e-5
    again = ops.render_rays(rays, base, 2.0, 6.0, 128, white_bkg=True, precision="bf16", seed=42)
    assert torch.equal(again["rgb"], whole["rgb"])            # idempotent / deterministic
    other = ops.render_rays(rays, base, 2.0, 6.0, 128, white_bkg=True, precision="bf16", seed=43)
    assert not torch.equal(other["rgb"], whole["rgb"])


def test_config1_64x64_32_coarse():
    """Config 1: 64x64, 32 coarse samples (the reference's render_image crashes at this size)."""
    H = W = 64
    pose = nerf_b200.pose_spherical(30.0, -30.0, 4.0)[:3, :]
    focal = nerf_b200.fov2Focal(FOV, (H, W))[0]
    rays = O.generate_rays(pose, H, W, focal)
    jitter = O.det_uniform((H * W, 32), 51, 0.0, 1.0)
    u = O.det_uniform((H * W, 129), 52, 0.0, 1.0)
    res = 4.0 / 32
    base = torch.linspace(2.0, 6.0 - res, 32)
    sp, sn = O.make_params("proposal", 1, "smooth"), O.make_params("nerf", 2, "smooth")
    ref = O.render_rays(sp, sn, rays, base, jitter, u, 2.0, 6.0, 128, white_bkg=False, resolution=res, softplus=True)
    net, prop = nets("smooth", "fp16x3")
    net._nb2_sync(); prop._nb2_sync()
    out = ops.render_rays(rays.to(DEV), base.to(DEV), 2.0, 6.0, 128, precision="fp16x3", jitter=jitter.to(DEV), u=u.to(DEV),
                          resolution=res, softplus=True)
    err = (out["rgb"].cpu() - ref["rgb"]).abs().max(dim=-1)[0]
    print("config1 max rgb err", float(err.max()), "frac > 1e-4", float((err > 1e-4).float().mean()))
    assert float((err > 1e-4).float().mean()) < 3e-2 and float(err.max()) < 5e-3


def test_reference_rng_mode_is_seed_reproducible():
    H = W = 100
    pose = nerf_b200.pose_spherical(10.0, -30.0, 4.0)[:3, :].to(DEV)
    focal = nerf_b200.fov2Focal(FOV, (H, W))[0]
    net, prop = nets("he", "bf16")
    torch.manual_seed(3)
    a = nerf_b200.render_image(net, prop, pose, (H, W), focal, 2.0, 6.0, 128, rng="reference")["rgb"]
    torch.manual_seed(3)
    b = nerf_b200.render_image(net, prop, pose, (H, W), focal, 2.0, 6.0, 128, rng="reference")["rgb"]
    assert torch.equal(a, b)
