"""GPU: the MLP kernels (CUDA-core fp32, tcgen05 bf16x3, tcgen05 bf16) against the reference outputs
(golden) and against the oracle run on the same inputs."""
import pytest
import torch

import nerf_b200
from nerf_b200 import _lib, ops
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"

# max |err| of the raw density relative to its scale (max |reference|, at least 50: the density head of the
# synthetic fields has std 25).  fp32: reordering noise only.  fp16x3: products exact to ~2^-22, the floor is
# the tensor core's truncating fp32 accumulation (~1e-6 per layer, same sign, 9 layers).  bf16x3: ~2^-16.  fp16 / bf16: 2^-11 / 2^-8 per operand through up to 9 layers.
REL_TOL = {"fp32": 1e-5, "fp16x3": 2e-5, "bf16x3": 1e-4, "fp16": 1e-2, "bf16": 6e-2}
RGB_TOL = {"fp32": 5e-6, "fp16x3": 2e-5, "bf16x3": 3e-5, "fp16": 5e-3, "bf16": 3e-2}
ALL_PREC = ["fp32", "fp16x3", "bf16x3", "fp16", "bf16"]


def scale(ref):
    return max(float(ref.abs().max()), 50.0)


def load(module, sd):
    module.load_state_dict({k: v.clone() for k, v in sd.items()})
    return module.to(DEV)


def test_tcgen05_building_blocks():
    """One UMMA sequence through the kernel's swizzle / descriptors / bulk copy / TMEM read-out."""
    g = torch.Generator().manual_seed(1)
    A = torch.randn(128, 64, generator=g).to(torch.bfloat16).to(DEV)
    B = torch.randn(128, 64, generator=g).to(torch.bfloat16).to(DEV)
    D = ops.selftest_umma(A, B)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().T
    assert float((D - ref).abs().max()) < 1e-3, float((D - ref).abs().max())


@pytest.mark.parametrize("precision", ALL_PREC)
@pytest.mark.parametrize("style", ["he", "smooth", "refinit"])
def test_mlp_forward_vs_reference(golden, gin, precision, style):
    prop = load(nerf_b200.ProposalNetwork(10, 256), O.make_params("proposal", 1, style))
    net = load(nerf_b200.MipNeRF(10, 4, 256), O.make_params("nerf", 2, style))
    prop.precision = net.precision = precision
    pts = gin["mlp_pts"].to(DEV)
    with torch.no_grad():
        p = prop.forward(pts[..., :3].contiguous()).cpu()
        n = net.forward(pts).cpu()
    gp, gn = golden[f"prop_fwd_{style}"], golden[f"nerf_fwd_{style}"]
    assert p.shape == gp.shape and n.shape == gn.shape
    tol = REL_TOL[precision]
    sc = 1.0 if style == "refinit" else 50.0   # refinit outputs are ~1e-4: compare on their own scale
    assert float((p - gp).abs().max()) <= tol * max(float(gp.abs().max()), sc), (float((p - gp).abs().max()), float(gp.abs().max()))
    assert float((n[..., 3] - gn[..., 3]).abs().max()) <= tol * max(float(gn[..., 3].abs().max()), sc)
    assert float((n[..., :3] - gn[..., :3]).abs().max()) <= RGB_TOL[precision]


@pytest.mark.parametrize("precision", ALL_PREC)
@pytest.mark.parametrize("n_points", [1, 127, 128, 129, 300 * 128 + 5])
def test_mlp_ragged_sizes(precision, n_points):
    sp, sn = O.make_params("proposal", 1, "he"), O.make_params("nerf", 2, "he")
    prop = load(nerf_b200.ProposalNetwork(10, 256), sp)
    net = load(nerf_b200.MipNeRF(10, 4, 256), sn)
    prop.precision = net.precision = precision
    pts = torch.cat((O.det_uniform((n_points, 3), 9, -2.0, 2.0), O.det_uniform((n_points, 3), 10, -1.0, 1.0)), -1).to(DEV)
    with torch.no_grad():
        p = prop.forward(pts[None, :, :3].contiguous())[0]
        n = net.forward(pts[None])[0]
    spd, snd = O.params_to(sp, DEV), O.params_to(sn, DEV)
    rp, rn = O.proposal_forward(spd, pts[:, :3]), O.nerf_forward(snd, pts)
    tol = REL_TOL[precision]
    assert float((p - rp).abs().max()) <= tol * scale(rp)
    assert float((n[:, 3] - rn[:, 3]).abs().max()) <= tol * scale(rn[:, 3])
    assert float((n[:, :3] - rn[:, :3]).abs().max()) <= RGB_TOL[precision]


def test_engine_error_is_at_the_reference_own_fp32_noise_level():
    """fp16x3 / bf16x3 vs an fp64 evaluation, next to the reference's own fp32 path vs fp64 (informational bound)."""
    sn = O.make_params("nerf", 2, "he")
    net = load(nerf_b200.MipNeRF(10, 4, 256), sn)
    pts = torch.cat((O.det_uniform((4096, 3), 9, -2.0, 2.0), O.det_uniform((4096, 3), 10, -1.0, 1.0)), -1)
    ref64 = torch.from_numpy(O.np_forward("nerf", sn, pts.numpy()))
    ref32 = O.nerf_forward(sn, pts).double()
    errs = {}
    for precision in ("fp32", "fp16x3", "bf16x3"):
        net.precision = precision
        with torch.no_grad():
            out = net.forward(pts[None].to(DEV))[0].cpu().double()
        errs[precision] = float((out[:, :3] - ref64[:, :3]).abs().max())
    e_ref = float((ref32[:, :3] - ref64[:, :3]).abs().max())
    print("rgb max|err| vs fp64: reference fp32", e_ref, "engine", errs)
    assert errs["fp32"] <= 5e-6 and errs["fp16x3"] <= 2e-5 and errs["bf16x3"] <= 3e-5


def test_repack_after_parameter_update():
    sn = O.make_params("nerf", 2, "he")
    net = load(nerf_b200.MipNeRF(10, 4, 256), sn)
    net.precision = "fp32"
    pts = torch.cat((O.det_uniform((256, 3), 9, -2.0, 2.0), O.det_uniform((256, 3), 10, -1.0, 1.0)), -1).to(DEV)[None]
    with torch.no_grad():
        a = net.forward(pts).clone()
        slot = net._nb2_sync()
        assert slot >= 2                         # every module instance owns its packed-network slot
        v0 = _lib.load().nb2_weights_version(_lib.handle(), slot)
        net.rgb_layer[2].bias.add_(0.5)          # in-place update, like an optimizer step
        b = net.forward(pts)
        v1 = _lib.load().nb2_weights_version(_lib.handle(), slot)
    assert v1 == v0 + 1
    assert float((a[..., :3] - b[..., :3]).abs().max()) > 1e-3 and torch.equal(a[..., 3], b[..., 3])


def test_unpacked_network_is_an_error():
    lib = _lib.load()
    import ctypes
    h = ctypes.c_void_p()
    _lib.check(lib.nb2_create(ctypes.byref(h), torch.cuda.current_device()))
    x = torch.zeros(128, 6, device=DEV)
    out = torch.zeros(128, 4, device=DEV)
    rc = lib.nb2_mlp_forward(h, _lib.NET_NERF, _lib.PREC_BF16, _lib.ptr(x), 6, 128, _lib.ptr(out), _lib.stream_ptr())
    assert rc == -3 and b"not been packed" in lib.nb2_last_error()
    lib.nb2_destroy(h)


def test_two_models_coexist_on_one_handle():
    """Train / eval / EMA copies: alternating forwards of two module instances do not re-pack each other."""
    a = load(nerf_b200.MipNeRF(10, 4, 256), O.make_params("nerf", 2, "he"))
    b = load(nerf_b200.MipNeRF(10, 4, 256), O.make_params("nerf", 3, "he"))
    a.precision = b.precision = "bf16"
    pts = torch.cat((O.det_uniform((512, 3), 9, -2.0, 2.0), O.det_uniform((512, 3), 10, -1.0, 1.0)), -1).to(DEV)[None]
    lib, h = _lib.load(), _lib.handle()
    with torch.no_grad():
        ya, yb = a.forward(pts).clone(), b.forward(pts).clone()
        sa, sb = a._nb2_sync(), b._nb2_sync()
        assert sa != sb
        va, vb = lib.nb2_weights_version(h, sa), lib.nb2_weights_version(h, sb)
        for _ in range(3):
            assert torch.equal(a.forward(pts), ya) and torch.equal(b.forward(pts), yb)
        assert (lib.nb2_weights_version(h, sa), lib.nb2_weights_version(h, sb)) == (va, vb)
    assert not torch.equal(ya, yb)
    del b
    import gc
    gc.collect()                                  # the slot goes back to the handle with the module
    c = load(nerf_b200.MipNeRF(10, 4, 256), O.make_params("nerf", 4, "he"))
    c.precision = "bf16"
    with torch.no_grad():
        c.forward(pts)
    assert c._nb2_sync() <= sb


def test_grad_mode_selects_the_differentiable_engine():
    """Parameters require grad and grad mode is on (every training call of train.py): forward records a graph through the
    layer-wise engine; under torch.no_grad() the fused inference kernel runs.  Both agree to the engines' precisions."""
    net = load(nerf_b200.MipNeRF(10, 4, 256), O.make_params("nerf", 2, "smooth"))
    pts = torch.cat((O.det_uniform((256, 3), 9, -2.0, 2.0), O.det_uniform((256, 3), 10, -1.0, 1.0)), -1).to(DEV)[None]
    a = net.forward(pts)
    assert a.requires_grad and a.grad_fn is not None
    with torch.no_grad():
        b = net.forward(pts)
    assert not b.requires_grad
    assert float((a.detach()[..., :3] - b[..., :3]).abs().max()) < 1e-4


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process():
    """Per-device launch state lives in the handle: a second GPU in the same process launches with its own shared-memory
    opt-in and cluster occupancy, on its own stream, without touching the caller's current device."""
    sn = O.make_params("nerf", 2, "he")
    pts = torch.cat((O.det_uniform((1000, 3), 9, -2.0, 2.0), O.det_uniform((1000, 3), 10, -1.0, 1.0)), -1)[None]
    outs = []
    for d in (0, 1):
        net = nerf_b200.MipNeRF(10, 4, 256)
        net.load_state_dict(sn)
        net = net.to(f"cuda:{d}")
        net.precision = "fp16x3"
        with torch.no_grad():
            outs.append(net.forward(pts.to(f"cuda:{d}")).cpu())
        assert torch.cuda.current_device() == 0
    assert torch.equal(outs[0], outs[1])


def test_fp16_modes_saturate_instead_of_overflowing():
    """Hidden activations beyond the fp16 range (65504): the fp16 packs saturate (cvt.rn.satfinite), so the fp16x3 / fp16
    modes return finite values instead of inf - inf = NaN rows; bf16x3 (fp32's exponent range) stays accurate."""
    sn = O.make_params("nerf", 2, "he")
    sn = {k: v.clone() for k, v in sn.items()}
    sn["lin_block1.0.weight"] *= 3.0e4                      # first hidden layer ~1e5
    sn["lin_block1.2.weight"] *= 1.0 / 3.0e4                # ... rescaled back by the next layer
    net = load(nerf_b200.MipNeRF(10, 4, 256), sn)
    pts = torch.cat((O.det_uniform((1024, 3), 9, -2.0, 2.0), O.det_uniform((1024, 3), 10, -1.0, 1.0)), -1).to(DEV)
    ref = O.nerf_forward(O.params_to(sn, DEV), pts)
    h1 = torch.relu(torch.nn.functional.linear(torch.cat((pts[:, :3], O.positional_encoding(pts[:, :3], 10)), -1), sn["lin_block1.0.weight"].to(DEV),
                                               sn["lin_block1.0.bias"].to(DEV)))
    assert float(h1.max()) > 65504.0
    with torch.no_grad():
        for precision in ("fp16x3", "fp16"):
            net.precision = precision
            assert bool(torch.isfinite(net.forward(pts[None])).all()), precision
        net.precision = "bf16x3"
        out = net.forward(pts[None])[0]
    assert float((out[:, :3] - ref[:, :3]).abs().max()) <= 1e-3
