"""CPU: the oracle (oracle/nerf_oracle.py) against outputs of the unmodified reference.

Tolerances: the oracle uses the same PyTorch ops as the reference, so on the machine that produced
the fixtures it is bit-identical; a few ulp are allowed because the GPU box's host CPU may pick
different SIMD kernels for exp/sin/sum.
"""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from tests.golden.make_golden import render_case

TOL = dict(rtol=2e-6, atol=2e-6)


def close(a, b, **kw):
    kw = {**TOL, **kw}
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.allclose(a, b, **kw), float((a - b).abs().max())


def test_positional_encoding(golden, gin):
    close(O.positional_encoding(gin["pe_x"], 10), golden["pe"], atol=1e-6)
    close(O.positional_encoding(gin["pe_x3"], 4), golden["pe3"], atol=1e-6)


def test_weights_and_blur(golden, gin):
    w = O.weights_from_sigma(gin["sigma"], gin["z"], gin["dirs"])
    close(w, golden["weights"])
    close(O.weights_from_sigma(gin["sigma"], gin["z"]), golden["weights_nodir"])
    close(O.max_blur(golden["weights"], 0.01), golden["blur"], atol=1e-7)
    # edge rows: empty ray -> zero weights; opaque first sample -> all weight on sample 0
    assert float(golden["weights"][0].abs().max()) == 0.0
    assert abs(float(golden["weights"][2, 0]) - 1.0) < 1e-6


def test_search_is_exact_given_cdf(golden, gin):
    """Stage (a): identical (cdf, u) -> identical indices (pure comparisons)."""
    mids = 0.5 * (gin["z"][:, 1:] + gin["z"][:, :-1])
    s, below, above = O.invert_cdf(golden["cdf"], mids, gin["u"])
    assert torch.equal(below, golden["pdf_below"])
    assert torch.equal(above, golden["pdf_above"])
    close(s, golden["pdf_samples"], atol=1e-6)


def _index_agreement(mine, ref):
    return float((mine == ref).float().mean())


def test_sample_pdf_and_inverse_sample(golden, gin):
    mids = 0.5 * (gin["z"][:, 1:] + gin["z"][:, :-1])
    w = golden["blur"]
    for torch_sum in (True, False):
        s, below, above = O.sample_pdf(mids, w[:, 1:-1], gin["u"], torch_sum=torch_sum)
        # documented-order cdf may differ from ATen's fp32 cascade sum by 1 ulp -> rare index flips
        assert _index_agreement(below, golden["pdf_below"]) >= 0.995
        bad = below != golden["pdf_below"]
        ok = ~bad
        close(s[ok], golden["pdf_samples"][ok], atol=2e-6)
        z, b = O.inverse_sample(w, gin["z"], gin["u"], sort=True, torch_sum=torch_sum)
        assert bool((z[:, 1:] >= z[:, :-1]).all())
        assert float((z - golden["inv_z"]).abs().max()) < 0.07   # a flipped index moves one z by < bin width
        assert float((z - golden["inv_z"]).abs().median()) < 1e-6
    zu = O.inverse_sample(w, gin["z"], gin["u"], sort=False, torch_sum=True)[0]
    assert float((zu - golden["inv_z_unsorted"]).abs().median()) < 1e-6


def test_points_merge_composite(golden, gin):
    close(O.length2pts(gin["rays"], gin["z_fine"]), golden["l2p"], atol=1e-6)
    pts, z = O.coarse_fine_merge(gin["rays"], gin["z"], golden["inv_z"])
    close(z, golden["merge_z"], atol=1e-6)
    close(pts, golden["merge_pts"], atol=1e-5)
    c = O.composite(gin["rgbo"], gin["z_fine"], gin["dirs"], white_bkg=True, near_far=(2.0, 6.0))
    close(c["rgb"], golden["comp_rgb"], atol=1e-6)
    close(c["weights"], golden["comp_w"], atol=1e-6)
    close(c["depth"], golden["comp_depth"], atol=1e-5)
    close(O.composite(gin["rgbo"], gin["z_fine"], gin["dirs"])["rgb"], golden["comp_rgb_black"], atol=1e-6)


def test_ipe(golden, gin):
    f, mu, mu_t = O.ipe_feature(gin["ipe_z"], gin["ipe_rays"], 10, 0.01)
    close(f, golden["ipe_feat"], atol=2e-6)
    close(mu, golden["ipe_mu"], atol=1e-6)
    close(mu_t, golden["ipe_mu_t"], atol=1e-6)


def test_mlps(golden, gin):
    for style in ("he", "smooth", "refinit"):
        sp, sn = O.make_params("proposal", 1, style), O.make_params("nerf", 2, style)
        p = O.proposal_forward(sp, gin["mlp_pts"][..., :3])
        n = O.nerf_forward(sn, gin["mlp_pts"])
        close(p, golden[f"prop_fwd_{style}"], rtol=1e-5, atol=1e-5 * float(golden[f"prop_fwd_{style}"].abs().max()))
        close(n, golden[f"nerf_fwd_{style}"], rtol=1e-5, atol=1e-5 * float(golden[f"nerf_fwd_{style}"].abs().max()))
    # the 'he' field is non-degenerate: densities of both signs, colours away from 0.5
    g = golden["nerf_fwd_he"]
    assert float(g[..., 3].max()) > 1.0 and float(g[..., 3].min()) < -1.0, (float(g[..., 3].max()), float(g[..., 3].min()))
    assert float(g[..., :3].std()) > 0.05


def test_render_image_tile(golden):
    """The reference's render_image on one 50x50 tile == oracle.render_rays on the same rays/uniforms."""
    H = W = 50
    pose, jitter, u, focal = render_case(H, W)
    rays = O.generate_rays(pose, H, W, focal)
    base_z = torch.linspace(2.0, 6.0, 64)
    for style in ("he", "smooth", "refinit"):
        sp, sn = O.make_params("proposal", 1, style), O.make_params("nerf", 2, style)
        out = O.render_rays(sp, sn, rays, base_z, jitter, u, 2.0, 6.0, 128, white_bkg=True, torch_sum=True)
        rgb = out["rgb"].view(H, W, 3).permute(2, 0, 1)
        err = (rgb - golden[f"img_rgb_{style}"]).abs()
        # same ops on the same machine: (near) bit-identical; on another host CPU the chaotic 'he' field
        # amplifies ulp-level exp/sin differences (see test_he_field_is_ill_conditioned)
        tol = 2e-3 if style == "he" else 1e-4
        assert float(err.max()) < tol, float(err.max())
        derr = (out["depth"].view(H, W) - golden[f"img_depth_{style}"]).abs()
        assert float(derr.max()) < tol, float(derr.max())


def test_he_field_is_ill_conditioned():
    """Why end-to-end 1e-4 parity is asserted on the band-limited field: on the He-init field a ONE-ulp change
    of the fine depths already moves the reference's own output by more than 1e-4."""
    H = W = 50
    pose, jitter, u, focal = render_case(H, W)
    rays = O.generate_rays(pose, H, W, focal)[::4]
    jitter, u = jitter[::4], u[::4]
    base_z = torch.linspace(2.0, 6.0, 64)
    moved = {}
    for style in ("he", "smooth"):
        sp, sn = O.make_params("proposal", 1, style), O.make_params("nerf", 2, style)
        z = O.render_rays(sp, sn, rays, base_z, jitter, u, 2.0, 6.0, 128, white_bkg=True)["z_fine"]
        z1 = torch.nextafter(z, torch.full_like(z, 10.0))

        def fine(zz):
            return O.composite(O.nerf_forward(sn, O.length2pts(rays, zz)), zz, rays[:, 3:], True, (2.0, 6.0))["rgb"]

        moved[style] = float((fine(z) - fine(z1)).abs().max())
    assert moved["he"] > 1e-4 and moved["smooth"] < 2e-5, moved


def test_config1_trainer_composition(golden):
    """Config 1 (64x64, 32 coarse samples): oracle vs the trainer-style composition of the reference."""
    sp, sn = O.make_params("proposal", 1, "he"), O.make_params("nerf", 2, "he")
    rays, lengths = golden["c1_rays"], golden["c1_lengths"]
    pts = rays[:, None, :3] + rays[:, None, 3:] * lengths[:, :, None]
    density = torch.nn.functional.softplus(O.proposal_forward(sp, pts))
    close(density, golden["c1_density"], rtol=1e-5, atol=1e-4)
    w = O.max_blur(O.weights_from_sigma(density, lengths, rays[:, 3:]), 0.01)
    u = O.det_uniform((256, 129), 42, 0.0, 1.0)
    fine, below = O.inverse_sample(w, lengths, u, sort=True, torch_sum=True)
    assert _index_agreement(below, golden["c1_below"]) >= 0.995
    fine = fine[..., :-1]
    rgbo = O.nerf_forward(sn, O.length2pts(rays, fine))
    rgb = O.composite(rgbo, fine, rays[:, 3:])["rgb"]
    assert float((rgb - golden["c1_rgb"]).abs().max()) < 1e-4


def test_philox_known_answers():
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors)."""
    import numpy as np

    def raw(ctr, key):
        c = [np.array([[x]], dtype=np.uint64) for x in ctr]
        k0, k1 = np.uint64(key[0]), np.uint64(key[1])
        M0, M1, W0, W1, m32 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), np.uint64(0xFFFFFFFF)
        for _ in range(10):
            p0, p1 = M0 * c[0], M1 * c[2]
            c = [(p1 >> np.uint64(32)) ^ c[1] ^ k0, p1 & m32, (p0 >> np.uint64(32)) ^ c[3] ^ k1, p0 & m32]
            k0, k1 = (k0 + W0) & m32, (k1 + W1) & m32
        return [int(x[0, 0]) for x in c]

    assert raw([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert raw([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert raw([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    # and the wrapper the CUDA kernels mirror: ray 0, samples 0..3, stream 0, seed 0 -> the first vector
    u = O.philox_uniform(0, [0], 4, 0)[0]
    exp = torch.tensor([(x >> 8) / 16777216.0 for x in [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]], dtype=torch.float32)
    assert torch.equal(u, exp)


def test_training_side_callers(golden, gin):
    """validSampler / getBounds / encoded_pt restatements against the reference."""
    from tests.golden.make_golden import inputs_train
    vs = inputs_train()
    pts, lengths, rgb, rays = O.valid_sampler(vs["rgbs"], vs["coords"], vs["cam_tf"], vs["indices"], vs["jitter"], 64, vs["focal"], 2.0, 6.0)
    assert torch.equal(lengths, golden["vs_len"]) and torch.equal(rgb, golden["vs_rgb"])
    close(rays, golden["vs_rays"], atol=1e-6)
    close(pts, golden["vs_pts"], atol=1e-5)
    close(O.get_bounds(golden["blur"], golden["inv_below"]), golden["bounds"], atol=1e-6)
    sp = O.make_params("proposal", 1, "smooth")
    out = O.proposal_forward(sp, golden["ipe_mu"], 10, encoded=golden["ipe_feat"])
    close(out, golden["prop_fwd_ipe"], rtol=1e-5, atol=1e-4)


def test_refnerf_helpers_oracle_matches_reference(golden_refnerf):
    """Ref-NeRF forward helpers: integrated directional encoding and linear_to_srgb, oracle vs the
    unmodified reference's outputs (ref_func.py:51-110, nerf_helper.py:50-56)."""
    from tests.golden.make_golden import inputs_refnerf
    g = inputs_refnerf()
    for deg in (1, 4, 5):
        out = O.ide(g["ide_dirs"], g["ide_kappa_inv"], deg)
        ref = golden_refnerf[f"ide_deg{deg}"]
        assert out.shape == ref.shape == (16, 16, 2 * sum(2 ** i + 1 for i in range(deg)))
        assert float((out - ref).abs().max()) <= 1e-6
    assert float((O.linear_to_srgb(g["srgb_lin"]) - golden_refnerf["srgb"]).abs().max()) <= 1e-6


def test_ide_tables_of_the_package_match_the_oracle():
    """Host logic of nerf_b200.ref_func.generate_ide_fn: (m, l) list and coefficient matrix (no GPU needed)."""
    import nerf_b200.ref_func as RF
    for deg in (1, 2, 4, 5):
        ml, mat = O.ide_tables(deg)
        arr = RF.get_ml_array(deg)
        assert [(int(m), int(l)) for m, l in arr.T] == ml
        l_max = 2 ** (deg - 1)
        m2 = torch.zeros(l_max + 1, arr.shape[1])
        for i, (m, l) in enumerate(arr.T):
            for k in range(l - m + 1):
                m2[k, i] = RF.sph_harm_coeff(int(l), int(m), k)
        assert torch.equal(m2, mat)
    with pytest.raises(ValueError):
        RF.generate_ide_fn(6)


def test_seeded_multi_tile_render_image(golden_round2):
    """The reference's render_image over several tiles with its OWN torch.rand draws under torch.manual_seed(2024)
    (tests/golden/make_golden.py round2): the oracle, fed the draws the engine's rng='reference' mode replays
    (nerf_b200.procedures._reference_rng_draws), reproduces the reference image -- including the rows the reference
    never renders when the height is not a whole number of tiles (70x100: rows 50..69 stay zero)."""
    import nerf_b200
    from nerf_b200 import procedures
    sp, sn = O.make_params("proposal", 1, "smooth"), O.make_params("nerf", 2, "smooth")
    for (H, W) in ((100, 100), (70, 100)):
        pose, _, _, _ = render_case(H, W)
        focal = float(nerf_b200.fov2Focal(0.6911112070083618, (H, W))[0]) if H == W else float(W / np.tan(0.5 * 0.6911112070083618))
        torch.manual_seed(2024)
        jit, u = procedures._reference_rng_draws((H, W), 64, 129)
        He = procedures.rendered_rows((H, W))
        rays = O.generate_rays(pose, H, W, focal)[:He * W]
        out = O.render_rays(sp, sn, rays, torch.linspace(2.0, 6.0, 64), jit, u, 2.0, 6.0, 128, white_bkg=True)
        img = torch.zeros(3, H, W)
        img[:, :He] = out["rgb"].view(He, W, 3).permute(2, 0, 1)
        g = golden_round2[f"seeded_rgb_{H}x{W}"]
        err = (img - g).abs().amax(0)
        assert bool((g[:, He:] == 0).all())
        # per-tile evaluation (2,500-row GEMMs) vs one 10,000-row batch: a few rays flip a cdf index
        assert float((err > 1e-4).float().mean()) < 5e-3 and float(err.median()) < 2e-6, (float(err.max()), float((err > 1e-4).float().mean()))


def test_training_step_gradients_match_the_reference(golden_round2):
    """The oracle's restated training step (train.py:164-199 + loss.backward()) against the UNMODIFIED reference's losses
    and parameter gradients on the same injected draws (tests/golden/make_golden.py round2)."""
    from tests.golden.make_golden import GRAD_HEAD, train_inputs
    ti = train_inputs()
    vs = ti["vs"]
    pts, lengths, rgb, rays = O.valid_sampler(vs["rgbs"], vs["coords"], vs["cam_tf"], ti["indices"], ti["jitter"], 64, vs["focal"], 2.0, 6.0)
    sp, sn = O.make_params("proposal", 1, "smooth"), O.make_params("nerf", 2, "smooth")
    out = O.train_step(sp, sn, pts, lengths, rgb, rays, ti["u"])
    g = golden_round2
    losses = torch.stack((out["loss"], out["img_loss"], out["prop_loss"]))
    assert float((losses - g["train_loss"]).abs().max()) <= 1e-5 * max(1.0, float(g["train_loss"].abs().max())), (losses, g["train_loss"])
    assert float((out["rendered"] - g["train_rendered"]).abs().max()) < 1e-5
    assert float((out["bounds"] - g["train_bounds"]).abs().max()) < 1e-5
    for tag, grads in (("prop", out["grad_prop"]), ("nerf", out["grad_nerf"])):
        for k, gr in grads.items():
            ref = g[f"grad_{tag}_{k}"]
            flat = gr.reshape(-1)
            got = torch.cat((flat.norm().reshape(1), flat.sum().reshape(1), flat[:GRAD_HEAD]))
            scale = max(float(ref[0]), 1e-12)
            assert float((got[0] - ref[0]).abs()) <= 1e-4 * scale, (tag, k, float(got[0]), float(ref[0]))
            assert float((got[2:] - ref[2:]).abs().max()) <= 1e-4 * scale, (tag, k)


def test_refnerf_forward_and_merge_indices(golden_round2):
    """The oracle's Ref-NeRF forward and coarseFineMerge index bookkeeping against the UNMODIFIED reference
    (nerf/ref_model.py:67-109, nerf/nerf_base.py:58-73; tests/golden/make_golden.py round2)."""
    from nerf_b200.ref_model import RefNeRF
    from tests.golden.make_golden import inputs_ops, refnerf_inputs
    g = golden_round2
    rn = RefNeRF(10, 4)
    sd = O.det_state_dict(rn, 7, gain=1.0)
    pts = refnerf_inputs()["pts"]
    rgbo, normal = O.refnerf_forward(sd, pts)
    assert float((rgbo - g["ref_fwd_rgbo"]).abs().max()) < 2e-5 and float((normal - g["ref_fwd_normal"]).abs().max()) < 2e-5
    rgbo_s, _ = O.refnerf_forward(sd, pts, use_srgb=True)
    assert float((rgbo_s - g["ref_fwd_rgbo_srgb"]).abs().max()) < 2e-5
    # merge with indices, restated: stable sort of cat(fine, coarse)
    go = inputs_ops()
    w = O.max_blur(O.weights_from_sigma(go["sigma"], go["z"], go["dirs"]), 0.01)
    zs, bs = O.inverse_sample(w, go["z"], go["u"], sort=True, torch_sum=True)
    zz, order = torch.sort(torch.cat((zs, go["z"]), dim=-1), dim=-1)
    assert float((zz[:, :-1] - g["merge_z2"]).abs().max()) < 1e-5


def ref_train_losses(forward, ti, normal_loss_func, bf_loss_func, img_loss_func):
    """train.py:176-199 (is_ref_model) on fixed samples; `forward(pos, dirs) -> (rgbo, normal)` with pos requiring grad."""
    import torch.nn.functional as F
    pos = ti["pos"].clone().requires_grad_(True)
    rgbo, normal = forward(pos, ti["dirs"])
    g, = torch.autograd.grad(rgbo[..., -1], pos, torch.ones_like(rgbo[..., -1]), retain_graph=True)
    dgrad = -g / torch.maximum(torch.full_like(g[..., :1], 1e-5), g.norm(dim=-1, keepdim=True))
    dens = F.softplus(rgbo[..., -1] + 0.5)
    w = O.weights_from_sigma(dens, ti["z"], None)          # train.py:182 passes density_act in mul_norm's place: no ||d|| scaling, relu
    rendered = torch.sum(w[:, :, None] * rgbo[..., :3], dim=-2)
    nl, bf, il = normal_loss_func(w, dgrad, normal), bf_loss_func(w, normal, ti["dirs"]), img_loss_func(rendered, ti["targets"])
    loss = il + 4e-4 * nl + 0.1 * bf
    return dict(pos=pos, dgrad=dgrad, weights=w, rendered=rendered, losses=torch.stack((loss, il, nl, bf)), loss=loss)


def test_refnerf_training_side_oracle_matches_reference(golden_ref_train):
    """RefNeRF.get_grad, WeightedNormalLoss, BackFaceLoss, coarse_grad_select and the parameter / position gradients of one
    is_ref_model loss: torch autograd over the oracle's forward against the UNMODIFIED reference (make_golden.py round3).
    This is what pins the checker of tests/test_gpu_i_refnerf.py's gradient tests."""
    import nerf_b200
    from tests.golden.make_golden import GRAD_HEAD, ref_train_inputs
    g = golden_ref_train
    ti = ref_train_inputs()
    rn = nerf_b200.RefNeRF(10, 4)
    sd = {k: v.clone().requires_grad_(True) for k, v in O.det_state_dict(rn, 7, gain=1.0).items()}
    r = ref_train_losses(lambda pos, dirs: O.refnerf_forward(sd, torch.cat((pos, dirs), -1)), ti,
                         nerf_b200.WeightedNormalLoss(True), nerf_b200.BackFaceLoss(), nerf_b200.SoftL1Loss())
    r["loss"].backward()
    assert float((r["dgrad"] - g["rt_density_grad"]).abs().max()) <= 2e-4          # unit vectors; fp32 summation order
    assert float((r["weights"] - g["rt_weights"]).abs().max()) <= 1e-5 and float((r["rendered"] - g["rt_rendered"]).abs().max()) <= 1e-5
    assert float(((r["losses"].detach() - g["rt_losses"]).abs() / g["rt_losses"].abs().clamp_min(1e-6)).max()) <= 1e-4
    assert float((r["pos"].grad - g["rt_pos_grad"]).norm() / g["rt_pos_grad"].norm()) <= 1e-3
    sel = nerf_b200.RefNeRF.coarse_grad_select(g["rt_density_grad"], ti["sort_inds"], 8)
    assert torch.equal(sel, g["rt_select"])
    worst = 0.0
    for k, p in sd.items():
        ref = g[f"rt_grad_{k}"]
        gk = p.grad.reshape(-1)
        worst = max(worst, abs(float(gk.norm()) - float(ref[0])) / max(float(ref[0]), 1e-12),
                    float((gk[:GRAD_HEAD] - ref[2:2 + GRAD_HEAD]).abs().max()) / max(float(ref[2:].abs().max()), 1e-12))
    print("Ref-NeRF training side, oracle vs reference: worst gradient deviation", worst)
    assert worst <= 2e-3
