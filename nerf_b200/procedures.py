"""Mirror of the reference's nerf/procedures.py: render_image, the hot path.

The reference walks 50x50-pixel tiles serially, ~1.1k PyTorch ops and two CPU-RNG -> GPU copies
per tile (reference nerf/procedures.py:60-90).  Here the whole image is one ray batch: rays are
generated on the device, and nb2_render_rays runs three kernels (fused sample+encode+proposal
MLP, resample, fused encode+NeRF MLP+composite) over all H*W rays.
"""
import os
from collections.abc import Iterable

import torch

from . import _lib, ops
from .addtional import ProposalNetwork
from .nerf_base import NeRF

POSSIBLE_PATCH_SIZE = [50, 40, 60, 30]
RENDER_COARSE_PNUM = 64


def get_patch_size(image_size):
    """Tile size of the reference (nerf/procedures.py:24-31): the first of [50, 40, 60, 30] that divides the WIDTH, and
    a (H // sz, W // sz) tile grid -- so when the height is not a multiple of sz the reference never renders the rows
    past (H // sz) * sz and leaves them at zero; render_image below reproduces that.  The one divergence: when no
    candidate divides the width the reference raises UnboundLocalError (e.g. 64x64); here that case is a single
    whole-image tile, (None, (1, 1))."""
    for patch_size in POSSIBLE_PATCH_SIZE:
        if image_size[1] % patch_size == 0:
            return patch_size, (image_size[0] // patch_size, image_size[1] // patch_size)
    return None, (1, 1)


def rendered_rows(image_size):
    """Rows of the image the reference's tile loop covers (all of them when no tile size applies)."""
    sz, patch_num = get_patch_size(image_size)
    return image_size[0] if sz is None else sz * patch_num[0]


def _reference_rng_draws(image_size, n_coarse, n_draw):
    """Replay the reference's CPU draws in its tile order and scatter them to raster order:
    per tile torch.rand((sz, sz, 64)) for the jitter (procedures.py:65) then torch.rand((sz*sz, 129))
    inside sample_pdf (utils.py:115).  Returns draws for the rendered rows only."""
    H, W = image_size
    sz, patch_num = get_patch_size(image_size)
    if sz is None:
        return torch.rand((H, W, n_coarse)).view(H * W, n_coarse), torch.rand((H * W, n_draw))
    He = sz * patch_num[0]
    jitter = torch.empty((He, W, n_coarse), dtype=torch.float32)
    u = torch.empty((He, W, n_draw), dtype=torch.float32)
    for k in range(patch_num[0]):
        for j in range(patch_num[1]):
            jitter[sz * k:sz * (k + 1), sz * j:sz * (j + 1)] = torch.rand((sz, sz, n_coarse))
            u[sz * k:sz * (k + 1), sz * j:sz * (j + 1)] = torch.rand((sz * sz, n_draw)).view(sz, sz, n_draw)
    return jitter.view(He * W, n_coarse), u.view(He * W, n_draw)


def render_image(
    network: NeRF, prop_net: ProposalNetwork, render_pose: torch.Tensor, image_size, focal,
    near: float, far: float, sample_num: int = 128, white_bkg: bool = False, render_depth=False, render_normal=False,
    *, precision=None, rng="philox", seed=None, jitter=None, u=None,
):
    """Same positional signature and return value as the reference (nerf/procedures.py:34-97):
    {"rgb": (3,H,W)[, "depth_img": (3,H,W)]} on render_pose.device.

    Keyword-only extensions: `precision` ('fp16x3' (default, fp32-faithful) | 'fp32' | 'bf16x3' | 'fp16' | 'bf16');
    `rng` = 'philox' (device counter-based RNG, seed from torch's CPU generator unless `seed` is given) or 'reference'
    (replays the reference's CPU torch.rand draws tile by tile, so the same torch.manual_seed gives the same samples
    as the reference); `jitter` (rows*W, 64) / `u` (rows*W, sample_num+1) inject the uniforms directly (raster order
    over the rendered rows).  Like the reference, rows past (H // tile) * tile stay zero (see get_patch_size).
    """
    if not isinstance(image_size, Iterable):
        image_size = (image_size, image_size)
    H, W = int(image_size[0]), int(image_size[1])
    from .ref_model import RefNeRF
    is_ref_model = type(network) == RefNeRF
    render_normal = bool(render_normal) and is_ref_model                        # procedures.py:41-42
    dev = render_pose.device
    if dev.type != "cuda":
        raise _lib.NB2Error("render_image: render_pose must live on a CUDA device (there is no CPU path)")
    if isinstance(focal, Iterable):
        fx, fy = float(focal[1]), float(focal[0])
    else:
        fx = fy = float(focal)
    if is_ref_model:
        return _render_image_ref(network, prop_net, render_pose, (H, W), (fx, fy), near, far, sample_num, white_bkg, render_depth,
                                 render_normal, precision, rng, seed, jitter, u)
    with torch.no_grad():
        nerf_id = network._nb2_sync()
        prop_id = prop_net._nb2_sync()
        He = rendered_rows((H, W))
        if He == 0:
            zero = torch.zeros((3, H, W), dtype=torch.float32, device=dev)
            return {"rgb": zero, "depth_img": zero.clone()} if render_depth else {"rgb": zero}
        rays = ops.generate_rays(render_pose, H, W, fx, fy, n_rays=He * W)
        base_z = torch.linspace(near, far, RENDER_COARSE_PNUM, device=dev)       # procedures.py:52
        resolution = (far - near) / sample_num                                    # procedures.py:59
        if jitter is None and u is None and rng == "reference":
            jitter, u = _reference_rng_draws((H, W), RENDER_COARSE_PNUM, sample_num + 1)
            jitter, u = jitter.to(dev, non_blocking=True), u.to(dev, non_blocking=True)
        elif rng not in ("philox", "reference"):
            raise ValueError(f"unknown rng mode {rng!r}")
        if seed is None:
            seed = ops._seed_from_torch() if (jitter is None or u is None) else 0
        prec = precision if precision is not None else (network.precision or prop_net.precision)
        out = ops.render_rays(rays, base_z, near, far, n_fine=sample_num, white_bkg=white_bkg, precision=prec,
                              jitter=jitter, u=u, seed=seed, resolution=resolution, prop_net_id=prop_id, nerf_net_id=nerf_id)
        def image(rows, channels):
            img = rows.view(He, W, channels).permute(2, 0, 1)
            if He == H:
                return img.contiguous()
            full = torch.zeros((channels, H, W), dtype=torch.float32, device=dev)   # the reference's never-rendered rows
            full[:, :He] = img
            return full
        result = {"rgb": image(out["rgb"], 3)}
        if render_depth:
            result["depth_img"] = image(out["depth"], 1).expand(3, H, W).contiguous()
    return result


def _render_image_ref(network, prop_net, render_pose, image_size, focal_xy, near, far, sample_num, white_bkg, render_depth, render_normal,
                      precision, rng, seed, jitter, u):
    """The Ref-NeRF branch of render_image (nerf/procedures.py:71-74,80-90): fused proposal kernel -> get_weights -> maxBlur ->
    inverseSample (129 kept) -> coarseFineMerge (193 sorted, last dropped = 192) -> RefNeRF.forward on the layer-wise
    engine -> softplus(density + 0.5) -> compositing with the normal image.  Staged launches (the fused fine kernel is
    MipNeRF's): ~40 kernels per image instead of 3."""
    H, W = image_size
    fx, fy = focal_xy
    dev = render_pose.device
    with torch.no_grad():
        He = rendered_rows((H, W))
        zero = torch.zeros((3, H, W), dtype=torch.float32, device=dev)
        if He == 0:
            res = {"rgb": zero}
            if render_depth:
                res["depth_img"] = zero.clone()
            if render_normal:
                res["normal_img"] = zero.clone()
            return res
        rays = ops.generate_rays(render_pose, H, W, fx, fy, n_rays=He * W)
        base_z = torch.linspace(near, far, RENDER_COARSE_PNUM, device=dev)
        resolution = (far - near) / sample_num
        if jitter is None and u is None and rng == "reference":
            jitter, u = _reference_rng_draws((H, W), RENDER_COARSE_PNUM, sample_num + 1)
            jitter, u = jitter.to(dev, non_blocking=True), u.to(dev, non_blocking=True)
        elif rng not in ("philox", "reference"):
            raise ValueError(f"unknown rng mode {rng!r}")
        if seed is None:
            seed = ops._seed_from_torch() if (jitter is None or u is None) else 0
        R = rays.shape[0]
        z_c, pts_c = ops.sample_coarse(rays, base_z, resolution, jitter=jitter, seed=seed)                  # procedures.py:65-66
        density = prop_net.forward(pts_c)                                                                   # :67
        w = ops.max_blur(ops.weights_from_sigma(density, z_c, rays[:, 3:].contiguous(), "relu"), 0.01)       # :68-69
        z_f, _ = ops.inverse_sample(w, z_c, sample_num + 1, sort=True, u=u, seed=seed)                       # :70
        pts, z = ops.coarse_fine_merge(rays, z_c, z_f)                                                       # :72
        P = z.shape[1]
        cam_dir = render_pose[:, -2].contiguous() if render_normal else None                                # :84
        rgbo, normal, ndot = network._forward_engine(pts.view(R * P, 6), pts.view(R * P, 6)[:, 3:6], cam_dir=cam_dir, shift_softplus=True)  # :73-74
        if render_normal:
            rgb, _, depth, _, n_img = ops.composite(rgbo.view(R, P, 4), z, rays[:, 3:].contiguous(), white_bkg=white_bkg,
                                                    near_far=(near, far) if render_depth else None, aux=ndot.view(R, P))
        else:
            rgb, _, depth, _ = ops.composite(rgbo.view(R, P, 4), z, rays[:, 3:].contiguous(), white_bkg=white_bkg,
                                             near_far=(near, far) if render_depth else None)

        def image(rows, channels):
            img = rows.view(He, W, channels).permute(2, 0, 1)
            if He == H:
                return img.contiguous()
            full = torch.zeros((channels, H, W), dtype=torch.float32, device=dev)
            full[:, :He] = img
            return full
        result = {"rgb": image(rgb, 3)}
        if render_depth:
            result["depth_img"] = image(depth, 1).expand(3, H, W).contiguous()
        if render_normal:
            result["normal_img"] = image((n_img + 1.0) * 0.5, 1).expand(3, H, W).contiguous()               # nerf_base.py:112
    return result


# ---- executables' call surface: argument parser and the render-only driver ---------------------------------
def get_parser():
    """The reference's command line (nerf/procedures.py:166-213): same flags, types and defaults, so train.py / ddp_train.py
    style drivers parse identically.  (tests/golden/reference_parser_defaults.json pins the defaults.)"""
    import argparse
    parser = argparse.ArgumentParser()
    ints = [("epochs", 2400), ("max_save", 3), ("sample_ray_num", 1024), ("coarse_sample_pnum", 64), ("fine_sample_pnum", 128),
            ("eval_time", 5), ("output_time", 20), ("center_crop_iter", 0), ("prop_net_width", 256), ("nerf_net_width", 256),
            ("decay_step", 100000), ("warmup_step", 500), ("ide_level", 4)]
    floats = [("near", 2.), ("far", 6.), ("center_crop_x", 0.5), ("center_crop_y", 0.5), ("img_scale", 0.5), ("scene_scale", 1.0),
              ("grad_clip", -0.01), ("pe_period_scale", 0.5), ("min_ratio", 0.01), ("decay_rate", 0.1), ("lr", 1.5e-4), ("bottle_neck_noise", 0.02)]
    strs = [("name", "model_1"), ("dataset_name", "lego"), ("opt_mode", "O1")]
    for name, default in ints:
        parser.add_argument("--" + name, type=int, default=default)
    for name, default in floats:
        parser.add_argument("--" + name, type=float, default=default)
    for name, default in strs:
        parser.add_argument("--" + name, type=str, default=default)
    for short, name in (("-d", "del_dir"), ("-l", "load"), ("-s", "use_scaler"), ("-b", "debug"), ("-v", "visualize"), ("-r", "do_render"),
                        ("-w", "white_bkg"), ("-t", "ref_nerf"), ("-u", "use_srgb"), ("-e", "eval_poses")):
        parser.add_argument(short, "--" + name, default=False, action="store_true")
    for name in ("render_depth", "render_normal", "prop_normal"):
        parser.add_argument("--" + name, default=False, action="store_true")
    return parser


def render_only(args, model_path: str, opt_level: str = "O1", dataset_root="../dataset/", output_root="./output/", max_frames=None):
    """The reference's render-only driver (nerf/procedures.py:99-164): load the two checkpoints, render the 120-pose orbit
    (or the test-set poses with PSNR against the ground truth when args.eval_poses) and save result_%03d.png.
    `args.use_scaler` selects the single-pass tensor precision (the analogue of the reference's AMP render); apex is not used."""
    from torchvision import transforms
    from torchvision.utils import save_image
    from .addtional import LossPSNR, SoftL1Loss
    from .dataset import AdaptiveResize, CustomDataSet
    from .mip_model import MipNeRF
    from .ref_model import RefNeRF
    from .utils import fov2Focal, pose_spherical
    eval_poses = args.eval_poses
    render_normal = args.render_normal & (not eval_poses)
    render_depth = args.render_depth & (not eval_poses)
    transform_funcs = transforms.Compose([AdaptiveResize(args.img_scale), transforms.ToTensor()])
    testset = CustomDataSet("%s%s/" % (dataset_root, args.dataset_name), transform_funcs, args.scene_scale, False, use_alpha=False)
    cam_fov_test, _ = testset.getCameraParam()
    r_c = testset.r_c()
    if eval_poses:
        all_poses = testset.tfs.cuda()
        loss_func, psnr_func = SoftL1Loss(), LossPSNR()
    else:
        all_poses = torch.stack([pose_spherical(angle, -30.0, 4.0) for angle in torch.linspace(-180, 180, 120 + 1)[:-1]], 0).cuda()
    test_focal = fov2Focal(cam_fov_test, r_c)
    if args.ref_nerf:
        mip_net = RefNeRF(10, args.ide_level, hidden_unit=args.nerf_net_width, perturb_bottle_neck_w=args.bottle_neck_noise, use_srgb=args.use_srgb).cuda()
    else:
        mip_net = MipNeRF(10, 4, hidden_unit=args.nerf_net_width).cuda()
    prop_net = ProposalNetwork(10, hidden_unit=args.prop_net_width).cuda()
    mip_net.loadFromFile(model_path + args.name + "_mip.pth", False)
    prop_net.loadFromFile(model_path + args.name + "_prop.pth", False)
    mip_net.eval()
    prop_net.eval()
    precision = "fp16m" if (args.use_scaler and not args.ref_nerf) else None
    output_dir = os.path.join(output_root, "given" if eval_poses else "sphere")
    os.makedirs(output_dir, exist_ok=True)
    psnrs = []
    with torch.no_grad():
        for i, pose in enumerate(all_poses):
            if max_frames is not None and i >= max_frames:
                break
            pose = pose.clone()
            pose[:3, -1] *= args.scene_scale
            result = render_image(mip_net, prop_net, pose[:3, :], r_c, test_focal, args.near, args.far, 128, white_bkg=args.white_bkg,
                                  render_normal=render_normal, render_depth=render_depth, precision=precision)
            if eval_poses:
                gt_img, _ = testset[i]
                gt_img = gt_img.cuda()
                loss = loss_func(result["rgb"], gt_img)
                psnr = psnr_func(loss)
                psnrs.append(float(psnr))
                print("Image loss:%.6f\tPSNR:%.4f" % (loss.item(), psnr.item()))
                result["gt_img"] = gt_img
            save_image(list(result.values()), os.path.join(output_dir, "result_%03d.png" % i), nrow=1 + render_depth + render_depth + eval_poses)
    return psnrs
