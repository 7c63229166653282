"""Staged GPU bring-up probe: each stage runs in its own process (a hang or fault in one cannot mask the
others) and dumps raw outputs under gpurun_out/ for offline analysis."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nerf_b200  # noqa: E402
from nerf_b200 import _lib, ops  # noqa: E402
from oracle import nerf_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
DEV = "cuda"


def load(module, sd):
    module.load_state_dict({k: v.clone() for k, v in sd.items()})
    return module.to(DEV)


def stage_selftest():
    g = torch.Generator().manual_seed(1)
    A = torch.randn(128, 64, generator=g).to(torch.bfloat16)
    B = torch.randn(128, 64, generator=g).to(torch.bfloat16)
    D = ops.selftest_umma(A.to(DEV), B.to(DEV))
    torch.cuda.synchronize()
    ref = A.float() @ B.float().T
    err = (D.cpu() - ref).abs()
    print("selftest max err", float(err.max()), "mean |ref|", float(ref.abs().mean()), "frac bad", float((err > 1e-2).float().mean()))
    np.savez(os.path.join(OUT, "selftest.npz"), A=A.float().numpy(), B=B.float().numpy(), D=D.cpu().numpy())


def stage_selftest_ts():
    """A-from-TMEM operand convention (tcgen05.st -> tcgen05.mma [d], [a], bdesc)."""
    lib, h = _lib.load(), _lib.handle()
    g = torch.Generator().manual_seed(1)
    A = torch.randn(128, 64, generator=g).to(torch.bfloat16).to(DEV)
    B = torch.randn(128, 64, generator=g).to(torch.bfloat16).to(DEV)
    scratch = torch.empty(16384, dtype=torch.uint8, device=DEV)
    D = torch.zeros(128, 128, device=DEV)
    _lib.check(lib.nb2_selftest_umma_ts(h, _lib.ptr(A), _lib.ptr(B), _lib.ptr(scratch), _lib.ptr(D), _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref = A.float() @ B.float().T
    err = (D - ref).abs()
    print("selftest_ts max err", float(err.max()), "mean |ref|", float(ref.abs().mean()), "frac bad", float((err > 1e-2).float().mean()))
    np.savez(os.path.join(OUT, "selftest_ts.npz"), A=A.float().cpu().numpy(), B=B.float().cpu().numpy(), D=D.cpu().numpy())


def stage_mlp(kind, precision, n=1000):
    sd = O.make_params(kind, 1 if kind == "proposal" else 2, "he")
    mod = load(nerf_b200.ProposalNetwork(10, 256) if kind == "proposal" else nerf_b200.MipNeRF(10, 4, 256), sd)
    mod.precision = precision
    pts = torch.cat((O.det_uniform((n, 3), 9, -2.0, 2.0), O.det_uniform((n, 3), 10, -1.0, 1.0)), -1)
    with torch.no_grad():
        t0 = time.time()
        out = mod.forward(pts[None].to(DEV) if kind == "nerf" else pts[None, :, :3].contiguous().to(DEV))[0]
        torch.cuda.synchronize()
        print("launch+sync s", time.time() - t0)
    ref = O.nerf_forward(sd, pts) if kind == "nerf" else O.proposal_forward(sd, pts[:, :3])
    err = (out.cpu() - ref).abs()
    print(kind, precision, "max err", float(err.max()), "max |ref|", float(ref.abs().max()), "nan", int(torch.isnan(out).sum()))
    if kind == "nerf":
        print("  rgb max err", float(err[:, :3].max()), "sigma max err", float(err[:, 3].max()))
    np.savez(os.path.join(OUT, f"mlp_{kind}_{precision}.npz"), out=out.cpu().numpy(), ref=ref.numpy())


def stage_time(precision, H=400):
    net = load(nerf_b200.MipNeRF(10, 4, 256), O.make_params("nerf", 2, "he"))
    prop = load(nerf_b200.ProposalNetwork(10, 256), O.make_params("proposal", 1, "he"))
    pose = nerf_b200.pose_spherical(30.0, -30.0, 4.0)[:3, :].to(DEV)
    focal = nerf_b200.fov2Focal(0.6911112070083618, (H, H))[0]
    for i in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        nerf_b200.render_image(net, prop, pose, (H, H), focal, 2.0, 6.0, 128, white_bkg=True, precision=precision, seed=1)
        torch.cuda.synchronize()
        dt = time.time() - t0
        print(f"render {H}x{H} {precision}: {dt * 1e3:.2f} ms  {H * H / dt / 1e6:.3f} Mrays/s")
    img = nerf_b200.render_image(net, prop, pose, (H, H), focal, 2.0, 6.0, 128, white_bkg=True, precision=precision, seed=1)["rgb"]
    print(f"VARIANT tmema={os.environ.get('NB2_TC_TMEMA','dflt')} nhalf={os.environ.get('NB2_TC_NHALF','dflt')} cluster={os.environ.get('NB2_TC_CLUSTER','dflt')} lockstep={os.environ.get('NB2_TC_LOCKSTEP','dflt')} {precision} "
          f"ms={dt * 1e3:.2f} checksum={float(img.double().sum()):.6f} nan={int(torch.isnan(img).sum())}")


def stage_ummabench():
    lib, h = _lib.load(), _lib.handle()
    g = torch.Generator().manual_seed(1)
    A = torch.randn(256, 64, generator=g).to(torch.bfloat16).to(DEV)
    B = torch.randn(256, 64, generator=g).to(torch.bfloat16).to(DEV)
    gsrc = torch.zeros(1 << 20, dtype=torch.uint8, device=DEV)
    ref = (A.float() @ B.float().T).cpu()

    def run(mode, iters, flags):
        D = torch.zeros(256, 256, device=DEV)
        cyc = torch.zeros(256, dtype=torch.int64, device=DEV)
        _lib.check(lib.nb2_debug_umma_bench(h, _lib.ptr(A), _lib.ptr(B), _lib.ptr(D), _lib.ptr(cyc), mode, iters, flags, _lib.ptr(gsrc), _lib.stream_ptr()))
        torch.cuda.synchronize()
        mr, nr = ((128, 128), (128, 256), (256, 256))[mode]
        err = float((D.cpu()[:mr, :nr] - ref[:mr, :nr]).abs().max())
        c = cyc.cpu()[:148]
        c = c[c > 0]
        print(f"UMMABENCH mode={mode} iters={iters} flags={flags:2d}: max|D-ref|={err:.2e} cycles/MMA median={float(c.float().median()):.1f} min={int(c.min())} max={int(c.max())}")

    for mode in (0, 1, 2):
        run(mode, 1, 0)
        run(mode, 400, 0)
    for flags in (1, 2, 3, 4, 8, 12, 16, 6, 7, 15, 31):
        run(0, 400, flags)
    for flags in (1, 2, 16):
        run(1, 400, flags)


def stage_roles(precision, H=400):
    """Per-role cycle accounting of the fine kernel (debug counters)."""
    import ctypes
    lib, h = _lib.load(), _lib.handle()
    _lib.check(lib.nb2_debug_tc_profile(h, None, 0))      # enable
    net = load(nerf_b200.MipNeRF(10, 4, 256), O.make_params("nerf", 2, "smooth"))
    prop = load(nerf_b200.ProposalNetwork(10, 256), O.make_params("proposal", 1, "smooth"))
    pose = nerf_b200.pose_spherical(30.0, -30.0, 4.0)[:3, :].to(DEV)
    focal = nerf_b200.fov2Focal(0.6911112070083618, (H, H))[0]
    for i in range(2):
        nerf_b200.render_image(net, prop, pose, (H, H), focal, 2.0, 6.0, 128, white_bkg=True, precision=precision, seed=1)
    buf = (ctypes.c_longlong * (148 * 16))()
    _lib.check(lib.nb2_debug_tc_profile(h, buf, 148))
    a = np.array(buf[:], dtype=np.int64).reshape(148, 16)
    names = ["str_wait_empty", "str_total", "ring_entries", "mma_wait_A", "mma_wait_W", "mma_total", "g0_encode", "g0_wait_acc",
             "g0_epi_hidden", "g0_epi_last", "g0_total", "iters", "layers", "g0_rgb_head_math", "g0_begin_next_tile", "-"]
    if os.environ.get("NB2_TC_NHALF") == "1":
        names[6:9] = ["nhalf:window_work", "nhalf:wait_acc_h0", "nhalf:drain_h0"]
    print(f"ROLES {precision} nhalf={os.environ.get('NB2_TC_NHALF','dflt')} cluster={os.environ.get('NB2_TC_CLUSTER','dflt')} lockstep={os.environ.get('NB2_TC_LOCKSTEP','dflt')} (median over CTAs, cycles; last launch = fine kernel)")
    lead = a[a[:, 5] > 0] if (a[:, 5] > 0).any() else a       # MMA counters exist on issuing CTAs only
    if precision in ("fp16x3", "bf16x3"):
        lead = a[0::2]                                        # split kernel: leaders are the even CTAs (odd rows carry the per-layer split)
    med = np.median(a, axis=0)
    med[3:6] = np.median(lead[:, 3:6], axis=0)
    for i, n in enumerate(names):
        print(f"   {n:16s} {med[i]:14.0f}   per-iter {med[i] / max(med[11], 1):12.0f}")
    if precision in ("fp16x3", "bf16x3"):
        ev = np.median(a[0::2][:, [13, 14]], axis=0) / max(med[11], 1)
        print(f"   rgb epilogue of the leader CTAs (per iter): accumulator read + head arithmetic {ev[0]:.0f}, hand-over of the next tile {ev[1]:.0f}")
    if precision in ("fp16x3", "bf16x3"):      # split kernel: the issuer's operand waits by layer, parked in the peer CTA's row
        peer = a[1::2][:, [0, 1, 2, 3, 4, 5, 13, 14, 15]]
        if (peer > 0).any():
            pm = np.median(peer, axis=0)
            print("   mma_wait_A by layer (per iter): " + "  ".join(f"L{l}={pm[l] / max(med[11], 1):.0f}" for l in range(9)))


def stage_hbm(R=160000):
    """Standalone HBM-bound ops at BASELINE config-2 sizes: achieved GB/s on the ALGORITHMIC bytes of SURVEY.md 8(d),
    CUDA events over 20 launches, L2 flushed (512 MB memset) before each timed launch."""
    import json
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    g = torch.Generator().manual_seed(3)
    pose = nerf_b200.pose_spherical(30.0, -30.0, 4.0)[:3, :].to(DEV)
    focal = float(nerf_b200.fov2Focal(0.6911112070083618, (400, 400))[0])
    rays = ops.generate_rays(pose, 400, 400, focal, focal)[:R].contiguous()
    base_z = torch.linspace(2.0, 6.0, 64, device=DEV)
    jitter = torch.rand(R, 64, generator=g).to(DEV)
    u = torch.rand(R, 129, generator=g).to(DEV)
    z, pts = ops.sample_coarse(rays, base_z, 4.0 / 128, jitter=jitter)
    sigma = (torch.rand(R, 64, generator=g) * 30.0).to(DEV)
    dirs = rays[:, 3:].contiguous()
    w = ops.weights_from_sigma(sigma, z, dirs)
    zf = ops.resample(sigma, z, rays, 129, u=u)
    rgbo = torch.rand(R, 128, 4, generator=g).to(DEV)
    x3 = pts.view(-1, 3)[: R * 8].contiguous()          # 1.28 M points
    cases = [
        ("generate_rays (a1)", lambda: ops.generate_rays(pose, 400, 400, focal, focal), 160000 * 24),
        ("sample_coarse (a3+a4)", lambda: ops.sample_coarse(rays, base_z, 4.0 / 128, jitter=jitter), R * (24 + 256 + 256 + 768)),
        ("posenc L=10 (a5)", lambda: ops.posenc(x3, 10), x3.shape[0] * (12 + 240)),
        ("ipe L=10 (a14)", lambda: ops.ipe(z, rays, 10, 0.01), R * 24 + R * 64 * 4 + R * 63 * (240 + 16)),
        ("weights_from_sigma (a7)", lambda: ops.weights_from_sigma(sigma, z, dirs), R * (256 + 256 + 12 + 256)),
        ("max_blur (a8)", lambda: ops.max_blur(w, 0.01), R * 512),
        ("inverse_sample sort (a9)", lambda: ops.inverse_sample(w, z, 129, sort=True, u=u), R * (256 + 256 + 516 + 516 + 1032)),
        ("resample fused (a7-a10)", lambda: ops.resample(sigma, z, rays, 129, u=u), R * (256 + 256 + 24 + 516 + 512)),
        ("length2pts (a10)", lambda: ops.length2pts(rays, zf), R * (24 + 512 + 128 * 24)),
        ("composite (a12)", lambda: ops.composite(rgbo, zf, dirs, white_bkg=True, near_far=(2.0, 6.0), want_weights=False), R * (128 * 20 + 12 + 16)),
    ]
    print(f"HBM ops at {R} rays; peak = {peaks['hbm_gbs']:.0f} GB/s (measured copy bandwidth)")
    for name, fn, nbytes in cases:
        for _ in range(3):
            fn()
        ts = []
        for _ in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        gbs = nbytes / (ms * 1e-3) / 1e9
        print(f"HBMOP {name:28s} {nbytes / 1e6:9.1f} MB  {ms * 1e3:9.1f} us  {gbs:8.1f} GB/s  {gbs / peaks['hbm_gbs']:.3f} of peak")


def stage_microbench():
    """TMEM load/store bandwidth and the L2 -> shared weight-stream ceiling (nb2_microbench.cu)."""
    import ctypes
    lib, h = _lib.load(), _lib.handle()
    out = torch.zeros(3 * 256, dtype=torch.int64, device=DEV)
    grid = ctypes.c_int(0)

    def run(kind, a0, a1, a2, a3=0, a4=0, src=None, n_chunks=0):
        out.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.nb2_debug_microbench(h, kind, a0, a1, a2, a3, a4, _lib.ptr(src) if src is not None else None, n_chunks,
                                            _lib.ptr(out), ctypes.byref(grid), _lib.stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        o = out.cpu().view(-1, 3)[:grid.value]
        cyc, byt = o[:, 0].double(), o[:, 1].double()
        return float((byt / cyc).median()), float((byt / cyc).min()), float(byt.sum()) / (e0.elapsed_time(e1) * 1e-3) / 1e12, grid.value

    for kind, name in ((0, "LDTM"), (1, "STTM")):
        for warps, batch in ((4, 1), (4, 2), (4, 4), (8, 1), (8, 2), (8, 4), (16, 1), (16, 2)):
            run(kind, warps, batch, 50)
            med, mn, tbs, g = run(kind, warps, batch, 400)
            print(f"MICRO {name} warps={warps:2d} in_flight={batch}: {med:7.1f} B/cycle/SM (min {mn:.1f}); 128x256 fp32 accumulator = {131072 / med:6.0f} cycles")
    src = torch.randint(0, 255, (256 * 16384,), dtype=torch.uint8, device=DEV)
    for cl, stages, skew, pairlike in ((1, 4, 0, 0), (1, 4, 0, 1), (1, 8, 0, 1), (1, 12, 0, 1), (1, 4, 7, 0), (1, 12, 7, 0),
                                       (2, 4, 0, 0), (2, 8, 0, 0), (4, 4, 0, 0), (4, 8, 0, 0), (4, 12, 0, 0), (8, 4, 0, 0), (8, 8, 0, 0), (8, 12, 0, 0)):
        run(2, cl, stages, 200, skew, pairlike, src, 256)
        med, mn, tbs, g = run(2, cl, stages, 4000, skew, pairlike, src, 256)
        print(f"MICRO L2STREAM cluster={cl} stages={stages:2d} skew={skew} pairlike={pairlike} grid={g}: {med:6.1f} B/cycle/SM (min {mn:.1f}), {tbs:5.2f} TB/s aggregate into shared memory")


if __name__ == "__main__":
    stage = sys.argv[1]
    print("== stage", stage, sys.argv[2:], "on", torch.cuda.get_device_name(0))
    if stage == "selftest_ts":
        stage_selftest_ts()
    elif stage == "selftest":
        stage_selftest()
    elif stage == "mlp":
        stage_mlp(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 1000)
    elif stage == "hbm":
        stage_hbm()
    elif stage == "microbench":
        stage_microbench()
    elif stage == "ummabench":
        stage_ummabench()
    elif stage == "roles":
        stage_roles(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 400)
    elif stage == "time":
        stage_time(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 400)
    print("== done", stage)
