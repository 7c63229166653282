#!/usr/bin/env python
"""bench.py — rays/s of the render hot path (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            the engine (libnerfb200.so)
  python bench.py --impl reference ...                     the reference algorithm on the host CPU cores
                                                           (the oracle port; /root/reference does not travel)
Workloads (synthetic orbit poses, random-init weights of the reference architectures, 64 coarse + 128 fine samples,
proposal 4x256 + vanilla NeRF 8x256, fp32-faithful arithmetic unless --precision says otherwise):
  N = 1   BASELINE configs[1]: one Lego-shaped 400x400 view per step.
  N > 1   BASELINE configs[4]: ONE 800x800 view per step, its 640,000 rays sharded contiguously across the N GPUs,
          rows gathered into the full image on every rank ("scaling": "strong").  The gather is fused into the
          compositing epilogue: finished rgb rows are stored straight into every GPU's image over NVLink (CUDA IPC peer
          mappings, nerf_b200/sharding.py:PeerImage); `--gather nccl` runs the all_gather baseline instead, and the
          line carries both numbers.  `--scaling weak` (one 400x400 view per GPU) is kept and reported as an extra key.

One JSON line on stdout (rank 0).  `value` = rays/s with inputs resident in HBM (CUDA events, max over ranks, L2 flushed
between timed steps); `e2e` = the same through the public API with the pose in pinned host memory and the full image
copied back to the host inside the timed region; `roofline` = the fused encode+MLP+composite kernel against the measured
bf16 tensor peak; `cpu_baseline` = the oracle on this box's host cores; `torch_cuda_baseline` = the reference algorithm
as PyTorch ops on the same B200 (fp32 with TF32 off, reference-style tile loop and whole-image chunks; fp16 autocast).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FOV = 0.6911112070083618
NEAR, FAR, N_COARSE, N_FINE = 2.0, 6.0, 64, 128
FLOP_PROP_PER_RAY = 27262976       # 2*(63*256 + 3*256^2 + 256) * 64
FLOP_NERF_PER_RAY = 135135232      # 2*(63*256 + 3*256^2 + 319*256 + 2*256^2 + 256^2 + 256 + 283*128 + 128*3) * 128
METRIC = "rays/sec (64c+128f samples)"
# dram__bytes_read.sum + dram__bytes_write.sum of ONE fine-kernel launch, per ray, from the ncu --set full captures
# summarised in profiles/ (algorithmic: 556 B/ray; the 2.4 MB of packed weights stay L2-resident)
NCU_FINE_TRAFFIC_BYTES_PER_RAY = {"fp16x3": 93.34e6 / 160000, "bf16x3": 93.34e6 / 160000, "fp16": 92.16e6 / 160000, "bf16": 92.16e6 / 160000}   # profiles/r02_ncu_tc{4_fp16x3,2_fp16}_fine.txt
FINE_KERNEL = {"fp16x3": "mlp_tc4_kernel (fine: encode + 8x256 MLP + composite; CTA-pair tcgen05, activations in TMEM)",
               "bf16x3": "mlp_tc4_kernel (fine: encode + 8x256 MLP + composite; CTA-pair tcgen05, activations in TMEM)",
               "fp16": "mlp_tc2_kernel (fine: encode + 8x256 MLP + composite; CTA-pair tcgen05, ping-pong tiles)",
               "bf16": "mlp_tc2_kernel (fine: encode + 8x256 MLP + composite; CTA-pair tcgen05, ping-pong tiles)",
               "fp32": "mlp_simt_kernel (fine: encode + 8x256 MLP on CUDA cores) + composite_kernel"}
DTYPE = {"fp16x3": "f32-faithful: fp16 hi+lo split operands, 3 tcgen05 MMAs per product, f32 accumulate", "bf16x3": "bf16 hi+lo split, f32 accumulate",
         "fp32": "f32 (CUDA cores)", "bf16": "bf16 operands, f32 accumulate", "fp16": "f16 operands, f32 accumulate"}
CPU_SAMPLE_HW = 100                # the CPU arms render a full 100x100 view (four reference tiles of 50x50)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"tflops": p.get("bf16_tflops_sustained", 1391.9), "hbm": p.get("hbm_gbs", 6549.1), "src": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"tflops": 1400.0, "hbm": 6650.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clock / throttle sampling DURING the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc = index, None
        self.path = tempfile.mktemp(suffix=".csv")

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_state_dicts():
    """Random-init weights of the reference architectures (band-limited 'smooth' field): nerf_b200/synthetic.py."""
    from nerf_b200 import synthetic as S
    return S.make_params("proposal", 1, "smooth"), S.make_params("nerf", 2, "smooth")


# ---- CPU arms (the oracle port of the reference algorithm; checker code, never on the engine's path) -------------------
def cpu_reference_render(steps, warmup, threads):
    """One step = the reference algorithm over a full 100x100 view in its own 50x50 tile order, PyTorch fp32 on the host."""
    import nerf_b200
    from oracle import nerf_oracle as O
    torch.set_num_threads(threads)
    sp, sn = synthetic_state_dicts()
    H = W = CPU_SAMPLE_HW
    pose = nerf_b200.pose_spherical(30.0, -30.0, 4.0)[:3, :]
    focal = nerf_b200.fov2Focal(FOV, (H, W))[0]
    rays = O.generate_rays(pose, H, W, focal).view(H, W, 6)
    base_z = torch.linspace(NEAR, FAR, N_COARSE)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            for k in range(H // 50):
                for j in range(W // 50):
                    r = rays[50 * k:50 * (k + 1), 50 * j:50 * (j + 1)].reshape(-1, 6)
                    jit, u = torch.rand(2500, N_COARSE), torch.rand(2500, N_FINE + 1)   # the reference's CPU draws, per tile
                    O.render_rays(sp, sn, r, base_z, jit, u, NEAR, FAR, N_FINE, white_bkg=True, chunk=2500)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    return H * W / t, t


def run_reference_arm(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    value, t = cpu_reference_render(args.steps, args.warmup, threads)
    n_rays = CPU_SAMPLE_HW * CPU_SAMPLE_HW
    sample = (f"{n_rays} rays per step: a full {CPU_SAMPLE_HW}x{CPU_SAMPLE_HW} view of the same scene in the reference's 50x50 tile order "
              f"(bounded sample of the 400x400 / 800x800 workload; rays/s normalises it), PyTorch {torch.__version__} CPU fp32, {threads} threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Lego-shaped orbit view, 64 coarse + 128 fine, proposal 4x256 + vanilla NeRF 8x256 (BASELINE configs[1]/[4]); bounded sample",
                   "rays_per_step": n_rays, "same_config": False},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def torch_cuda_baseline(dev, H, W):
    """The like-for-like "before": the reference algorithm as PyTorch ops on the SAME B200 (oracle port; the reference's own
    files do not travel to the GPU box).  fp32 with TF32 off in the reference's 50x50 tile loop with its per-tile CPU draws
    (nerf/procedures.py:60-90), fp32 in whole-image chunks with device-resident uniforms, and fp16 autocast
    (nerf/procedures.py:149-152)."""
    import nerf_b200
    from oracle import nerf_oracle as O
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sp, sn = synthetic_state_dicts()
    sp, sn = O.params_to(sp, dev), O.params_to(sn, dev)
    pose = nerf_b200.pose_spherical(30.0, -30.0, 4.0)[:3, :].to(dev)
    focal = nerf_b200.fov2Focal(FOV, (H, W))[0]
    base_z = torch.linspace(NEAR, FAR, N_COARSE, device=dev)
    n = H * W
    out = {}

    def tile_loop():
        rays = O.generate_rays(pose, H, W, focal).view(H, W, 6)
        for k in range(H // 50):
            for j in range(W // 50):
                r = rays[50 * k:50 * (k + 1), 50 * j:50 * (j + 1)].reshape(-1, 6)
                jit, u = torch.rand(2500, N_COARSE).to(dev), torch.rand(2500, N_FINE + 1).to(dev)
                O.render_rays(sp, sn, r, base_z, jit, u, NEAR, FAR, N_FINE, white_bkg=True, chunk=2500)

    jit_d, u_d = torch.rand(n, N_COARSE, device=dev), torch.rand(n, N_FINE + 1, device=dev)

    def chunked():
        O.render_rays(sp, sn, O.generate_rays(pose, H, W, focal), base_z, jit_d, u_d, NEAR, FAR, N_FINE, white_bkg=True, chunk=20000)

    def autocast():
        with torch.autocast("cuda", dtype=torch.float16):
            chunked()

    with torch.no_grad():
        for name, fn in (("fp32_tile_loop", tile_loop), ("fp32_chunked", chunked), ("fp16_autocast_chunked", autocast)):
            fn()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(2):
                fn()
            torch.cuda.synchronize(dev)
            t = (time.perf_counter() - t0) / 2
            out[name] = {"rays_per_s": n / t, "ms_per_image": 1e3 * t}
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    out["what"] = (f"oracle port of the reference algorithm, PyTorch {torch.__version__} CUDA ops on this GPU, one {H}x{W} view, TF32 off; "
                   "tile_loop = 50x50 tiles with per-tile CPU torch.rand + H2D as the reference does")
    return out


def train_step_leg(dev, precision="bf16x3", batches=(1024, 8192), steps=10, ddp=None):
    """One training step of the non-Ref model as the reference's trainer runs it (train.py:157-218): validSampler ->
    run() closure -> loss.backward() -> Adam step, on `R` rays x (64 coarse + 128 fine) samples, through the layer-wise
    tcgen05 engine (nerf_b200/train_engine.py).  R = 1024 is the reference's default --sample_ray_num; the larger batch
    shows the engine once the ~75 launches of a step stop being launch-bound.  Next to it: the same step as PyTorch ops on
    the same GPU (oracle port, fp32, TF32 off).
    ddp = (rank, world): the data-parallel step of ddp_train.py (every rank its own R rays; the one collective is the gradient
    all-reduce, ddp_train.py:98, here ONE flat NCCL all-reduce of 743,051 fp32 values for both networks); times are the max
    over ranks, rays/s the whole job's, and the replicas' parameters are checked to stay bit-identical."""
    import torch.nn.functional as F
    import nerf_b200
    from nerf_b200 import NeRF, ProposalNetwork, getBounds, inverseSample, maxBlurFilter
    from nerf_b200.train_engine import allreduce_gradients
    from oracle import nerf_oracle as O
    import torch.distributed as dist
    world = ddp[1] if ddp else 1
    if ddp:
        torch.manual_seed(1234 + ddp[0])               # every rank samples its own rays
    peaks = load_peaks()
    sd_prop, sd_nerf = synthetic_state_dicts()
    prop_net, mip_net = nerf_b200.ProposalNetwork(10, 256), nerf_b200.MipNeRF(10, 4, 256)
    prop_net.load_state_dict(sd_prop); mip_net.load_state_dict(sd_nerf)
    prop_net, mip_net = prop_net.to(dev), mip_net.to(dev)
    prop_net.train_precision = mip_net.train_precision = precision
    grad_vars = list(mip_net.parameters()) + list(prop_net.parameters())
    opt = torch.optim.Adam(params=grad_vars, lr=1.5e-4, betas=(0.9, 0.999))
    loss_func, prop_loss_func = nerf_b200.SoftL1Loss(), nerf_b200.ProposalLoss()
    Hh = Ww = 400
    g = torch.Generator().manual_seed(3)
    rgbs = torch.rand(Hh * Ww, 3, generator=g).to(dev)
    rows, cols = torch.meshgrid(torch.arange(Hh), torch.arange(Ww), indexing="ij")
    coords = torch.stack((cols - Ww // 2, Hh // 2 - rows), dim=-1).reshape(-1, 2).to(dev)
    cam_tf = nerf_b200.pose_spherical(30.0, -30.0, 4.0)[:3, :].contiguous().to(dev)
    focal = nerf_b200.fov2Focal(FOV, (Hh, Ww))
    passes = 3 if precision == "bf16x3" else 1
    # algorithmic HBM bytes per MLP row of the layer-wise engine: every Linear (in k, out n) reads X and writes Y in the
    # forward, reads dY + the relu mask and writes dX in the dgrad, reads dY and X in the wgrad; bf16 hi (+ lo) operands
    b16 = 2 * (2 if precision == "bf16x3" else 1)

    def layer_bytes(layers, first_has_dgrad=False):
        t = 0
        for i, (k, n) in enumerate(layers):
            t += b16 * k + b16 * n                      # forward
            t += b16 * n + b16 * k                      # wgrad
            if i > 0 or first_has_dgrad:
                t += b16 * n + 2 * k + b16 * k          # dgrad (+ mask)
        return t
    nerf_layers = [(64, 256), (256, 256), (256, 256), (256, 256), (320, 256), (256, 256), (256, 256), (256, 256), (256, 8), (288, 128), (128, 8)]
    prop_layers = [(64, 256), (256, 256), (256, 256), (256, 256), (256, 8)]
    bytes_per_ray = N_FINE * layer_bytes(nerf_layers) + N_COARSE * layer_bytes(prop_layers)
    out = {"precision": precision, "mma_passes_per_product": passes, "what": train_step_leg.__doc__.split("\n")[0]}
    for R in batches:
        def step(i):
            coarse_samples, coarse_lengths, rgb_targets, coarse_cam_rays = nerf_b200.validSampler(rgbs, coords, cam_tf, R, N_COARSE, focal, NEAR, FAR, True)
            density = F.softplus(prop_net.forward(coarse_samples))
            prop_weights = maxBlurFilter(ProposalNetwork.get_weights(density, coarse_lengths, coarse_cam_rays[:, 3:]), 0.01)
            fine_lengths, below_idxs = inverseSample(prop_weights, coarse_lengths, N_FINE + 1, sort=True)
            fine_lengths = fine_lengths[..., :-1]
            fine_rgbo = mip_net.forward(NeRF.length2pts(coarse_cam_rays, fine_lengths))
            fine_rendered, weights, _ = NeRF.render(fine_rgbo, fine_lengths, coarse_cam_rays[:, 3:])
            weight_bounds = getBounds(prop_weights, below_idxs)
            opt.zero_grad()
            loss = prop_loss_func(weight_bounds, weights.detach()) + loss_func(fine_rendered, rgb_targets)
            loss.backward()
            if ddp:
                allreduce_gradients([mip_net, prop_net])
            opt.step()
            return loss
        for i in range(3):
            step(i)
        torch.cuda.synchronize(dev)
        if ddp:
            dist.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t0 = time.perf_counter()
        for i in range(steps):
            ev[i][0].record()
            loss = step(10 + i)
            ev[i][1].record()
        torch.cuda.synchronize(dev)
        wall = (time.perf_counter() - t0) / steps
        ms = sum(a.elapsed_time(b) for a, b in ev) / steps
        flop = 3 * (FLOP_PROP_PER_RAY + FLOP_NERF_PER_RAY) * R
        if ddp:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
            flat = torch.cat([p.detach().reshape(-1) for p in grad_vars])
            lo, hi = flat.clone(), flat.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            out[f"rays_{R}_per_gpu"] = {
                "rays_per_step": R * world, "ms_per_step": ms, "host_ms_per_step": 1e3 * wall, "rays_per_s": R * world / (ms * 1e-3), "loss_rank0": float(loss),
                "allreduce_bytes_per_step": 4 * flat.numel(), "replica_param_max_abs_diff": float((hi - lo).abs().max()),
                "tensor_tflops": world * flop / (ms * 1e-3) / 1e12,
                "hbm_frac_per_gpu": bytes_per_ray * R / (ms * 1e-3) / 1e9 / peaks["hbm"]}
            continue
        # the reference algorithm's step as PyTorch ops on this GPU
        sp, sn = O.params_to(sd_prop, dev), O.params_to(sd_nerf, dev)
        cs, cl, rt, cr = nerf_b200.validSampler(rgbs, coords, cam_tf, R, N_COARSE, focal, NEAR, FAR, True)
        u = torch.rand(R, N_FINE + 1, device=dev)
        tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        O.train_step(sp, sn, cs, cl, rt, cr, u)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(3):
            O.train_step(sp, sn, cs, cl, rt, cr, u)
        torch.cuda.synchronize(dev)
        ref_ms = 1e3 * (time.perf_counter() - t0) / 3
        torch.backends.cuda.matmul.allow_tf32 = tf32
        out[f"rays_{R}"] = {
            "rays_per_step": R, "ms_per_step": ms, "host_ms_per_step": 1e3 * wall, "rays_per_s": R / (ms * 1e-3), "loss": float(loss),
            "roofline": {"bound": "hbm", "achieved": bytes_per_ray * R / (ms * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                         "frac": bytes_per_ray * R / (ms * 1e-3) / 1e9 / peaks["hbm"], "algorithmic_bytes_per_step": bytes_per_ray * R,
                         "algorithmic_flop_per_step": flop, "tensor_tflops": flop / (ms * 1e-3) / 1e12,
                         "frac_of_tensor_peak": flop / (ms * 1e-3) / 1e12 / peaks["tflops"],
                         "frac_of_tensor_peak_issued": passes * flop / (ms * 1e-3) / 1e12 / peaks["tflops"]},
            "torch_cuda_fp32_ms_per_step": ref_ms, "speedup_vs_torch_cuda_fp32": ref_ms / ms}
    return out


def config3_leg(dev, prop, net, ids, steps=5):
    """BASELINE configs[2]: Mip-NeRF-style integrated positional encoding feeding the proposal network, 400x400, single-pass
    tensor precision.  IPE is evaluated in the proposal kernel's producer (NB2_PROPOSAL_IPE); PSNR delta against the
    fp32-faithful IPE render at a 30 dB synthetic ground truth (north_star: within 0.05 dB)."""
    import nerf_b200
    from nerf_b200 import ops
    H = W = 400
    pose = nerf_b200.pose_spherical(30.0, -30.0, 4.0)[:3, :].to(dev)
    focal = float(nerf_b200.fov2Focal(FOV, (H, W))[0])
    radius = 2.0 / (focal * 12 ** 0.5)
    base_z = torch.linspace(NEAR, FAR, N_COARSE, device=dev)
    rays = ops.generate_rays(pose, H, W, focal, focal)
    ref = ops.render_rays(rays, base_z, NEAR, FAR, N_FINE, white_bkg=True, precision="fp16x3", seed=5, ipe_radius=radius, **ids)["rgb"].clone()
    gt = ref + 0.0316 * torch.randn(ref.shape, generator=torch.Generator().manual_seed(1)).to(dev)
    psnr = lambda a: -10.0 * math.log10(float(((a - gt) ** 2).mean()))
    out = {"workload": "400x400 view, 64 coarse + 128 fine, integrated positional encoding (cone radius 2/sqrt(12) pixel) in the proposal kernel's producer",
           "what": config3_leg.__doc__.split("\n")[0]}
    for mode in ("fp16", "bf16", "fp16m", "bf16m", "fp16x3"):
        res = None
        for _ in range(2):
            res = ops.render_rays(rays, base_z, NEAR, FAR, N_FINE, white_bkg=True, precision=mode, seed=5, ipe_radius=radius, out=res, **ids)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            res = ops.render_rays(rays, base_z, NEAR, FAR, N_FINE, white_bkg=True, precision=mode, seed=5, ipe_radius=radius, out=res,
                                  workspace=res["_workspace"], **ids)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        out[mode] = {"rays_per_s": H * W / (ms * 1e-3), "ms_per_step": ms, "psnr_delta_db_at_30dB_gt": psnr(res["rgb"]) - psnr(ref)}
    return out


def config4_leg(dev, prop, steps=5):
    """BASELINE configs[3]: Ref-NeRF forward on 512-ray batches (64 coarse + 129 fine merged to 192 samples, proposal + IDE):
    fused proposal kernel -> resample -> coarseFineMerge -> RefNeRF (17 layers on the layer-wise tcgen05 engine, IDE kernel)
    -> compositing with depth and normal outputs; `train_512`: the reference's Ref-NeRF training step (refnerf_train_leg)."""
    import nerf_b200
    from nerf_b200 import synthetic
    rn = nerf_b200.RefNeRF(10, 4)
    rn.load_state_dict(synthetic.det_state_dict(rn, 7, gain=1.0))
    rn = rn.to(dev).eval()
    pose = nerf_b200.pose_spherical(30.0, -30.0, 4.0)[:3, :].to(dev)
    out = {"what": config4_leg.__doc__.split("\n")[0]}
    for (Hh, Ww) in ((16, 32), (100, 100)):                     # 512 rays (the config's batch) and a 10,000-ray image
        focal = float(nerf_b200.fov2Focal(FOV, (Ww, Ww))[0])
        for _ in range(2):
            nerf_b200.render_image(rn, prop, pose, (Hh, Ww), focal, NEAR, FAR, N_FINE, white_bkg=True, render_depth=True, render_normal=True, seed=3)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            nerf_b200.render_image(rn, prop, pose, (Hh, Ww), focal, NEAR, FAR, N_FINE, white_bkg=True, render_depth=True, render_normal=True, seed=3)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        flop = Hh * Ww * (FLOP_PROP_PER_RAY + 192 * 2143232)
        out[f"rays_{Hh * Ww}"] = {"rays_per_s": Hh * Ww / (ms * 1e-3), "ms_per_step": ms, "host_ms_per_step": 1e3 * (time.perf_counter() - t0) / steps,
                                  "tensor_tflops": flop / (ms * 1e-3) / 1e12}
    with torch.enable_grad():      # the caller renders under no_grad
        out["train_512"] = refnerf_train_leg(dev, rn, prop)
    return out


def refnerf_train_leg(dev, rn, prop, R=512, steps=10):
    """The reference's training step with is_ref_model (train.py:164-218) on 512 rays x (64 coarse, 129 fine merged to 192):
    proposal network -> resample -> coarseFineMerge -> RefNeRF.forward(fine_pos, fine_dir) -> get_grad (density normals) ->
    render -> image + proposal + normal + back-face losses -> backward -> Adam.  Next to it the same step written with
    PyTorch fp32 ops on this GPU (the oracle's functions, TF32 off)."""
    import torch.nn.functional as F
    import nerf_b200
    from nerf_b200 import NeRF, ProposalNetwork, RefNeRF, getBounds, inverseSample, maxBlurFilter
    from oracle import nerf_oracle as O
    prop_eval = prop
    prop = nerf_b200.ProposalNetwork(10, 256)          # a copy: the optimizer steps must not touch the network the other legs render with
    prop.load_state_dict(prop_eval.state_dict())
    prop = prop.to(dev)
    rn.train()
    opt = torch.optim.Adam(list(rn.parameters()) + list(prop.parameters()), lr=1.5e-4)
    normal_loss_func, bf_loss_func = nerf_b200.WeightedNormalLoss(True), nerf_b200.BackFaceLoss()
    prop_loss_func, loss_func = nerf_b200.ProposalLoss(), nerf_b200.SoftL1Loss()
    Hh = Ww = 400
    g = torch.Generator().manual_seed(3)
    rgbs = torch.rand(Hh * Ww, 3, generator=g).to(dev)
    rows, cols = torch.meshgrid(torch.arange(Hh), torch.arange(Ww), indexing="ij")
    coords = torch.stack((cols - Ww // 2, Hh // 2 - rows), dim=-1).reshape(-1, 2).to(dev)
    cam_tf = nerf_b200.pose_spherical(30.0, -30.0, 4.0)[:3, :].contiguous().to(dev)
    focal = nerf_b200.fov2Focal(FOV, (Hh, Ww))
    keep = {}

    def step():
        cs, cl, rgb_targets, cr = nerf_b200.validSampler(rgbs, coords, cam_tf, R, N_COARSE, focal, NEAR, FAR, True)
        density = F.softplus(prop.forward(cs))
        prop_weights = maxBlurFilter(ProposalNetwork.get_weights(density, cl, cr[:, 3:]), 0.01)
        fine_lengths, below_idxs = inverseSample(prop_weights, cl, N_FINE + 1, sort=True)
        fine_samples, fine_lengths, below_idxs, _ = NeRF.coarseFineMerge(cr, cl, fine_lengths, below_idxs)
        fine_pos, fine_dir = fine_samples.split((3, 3), dim=-1)
        fine_pos.requires_grad = True
        fine_rgbo, pred_normal = rn.forward(fine_pos, fine_dir)
        density_grad = -RefNeRF.get_grad(fine_rgbo[..., -1], fine_pos)
        fine_rgbo[..., -1] = F.softplus(fine_rgbo[..., -1] + 0.5)
        fine_rendered, weights, _ = NeRF.render(fine_rgbo, fine_lengths, cr[:, 3:], rn.density_act)
        normal_loss = normal_loss_func(weights, density_grad, pred_normal)
        bf_loss = bf_loss_func(weights, pred_normal, fine_dir)
        weight_bounds = getBounds(prop_weights, below_idxs)
        opt.zero_grad()
        loss = prop_loss_func(weight_bounds, weights.detach()) + loss_func(fine_rendered, rgb_targets) + 4e-4 * normal_loss + 0.1 * bf_loss
        loss.backward()
        opt.step()
        keep.update(cs=cs.detach(), cl=cl, rt=rgb_targets, cr=cr, pos=fine_pos.detach(), dirs=fine_dir.detach(), fl=fine_lengths.detach(),
                    wb=weight_bounds.detach())
        return loss
    for _ in range(3):
        step()
    torch.cuda.synchronize(dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    t0 = time.perf_counter()
    for i in range(steps):
        ev[i][0].record()
        loss = step()
        ev[i][1].record()
    torch.cuda.synchronize(dev)
    wall = (time.perf_counter() - t0) / steps
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    # the same arithmetic as PyTorch ops: both MLPs forward + backward on the same samples (sampling / sorting excluded: in favour of the baseline)
    sd_r = {k: v.detach().clone().requires_grad_(True) for k, v in rn.state_dict().items()}
    sd_p = {k: v.detach().clone().requires_grad_(True) for k, v in prop.state_dict().items()}
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False

    def torch_step():
        dens = F.softplus(O.proposal_forward(sd_p, keep["cs"]))
        pw = O.max_blur(O.weights_from_sigma(dens, keep["cl"], keep["cr"][:, 3:]), 0.01)
        pos = keep["pos"].clone().requires_grad_(True)
        rgbo, normal = O.refnerf_forward(sd_r, torch.cat((pos, keep["dirs"]), -1))
        gd, = torch.autograd.grad(rgbo[..., -1], pos, torch.ones_like(rgbo[..., -1]), retain_graph=True)
        dg = -gd / torch.maximum(torch.full_like(gd[..., :1], 1e-5), gd.norm(dim=-1, keepdim=True))
        w = O.weights_from_sigma(F.softplus(rgbo[..., -1] + 0.5), keep["fl"], None)
        rendered = torch.sum(w[:, :, None] * rgbo[..., :3], dim=-2)
        l = (loss_func(rendered, keep["rt"]) + 4e-4 * normal_loss_func(w, dg, normal) + 0.1 * bf_loss_func(w, normal, keep["dirs"])
             + (pw.sum() * 0.0))
        l.backward()
        return l
    torch_step()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(3):
        torch_step()
    torch.cuda.synchronize(dev)
    ref_ms = 1e3 * (time.perf_counter() - t0) / 3
    torch.backends.cuda.matmul.allow_tf32 = tf32
    rn.eval()
    # forward + parameter backward twice (get_grad re-runs the backward plan) ~ 5 x forward FLOP of RefNeRF, 3 x of the proposal network
    flop = R * (3 * FLOP_PROP_PER_RAY + 5 * 192 * 2143232)
    return {"what": refnerf_train_leg.__doc__.split("\n")[0], "rays_per_step": R, "ms_per_step": ms, "host_ms_per_step": 1e3 * wall,
            "rays_per_s": R / (ms * 1e-3), "loss": float(loss), "tensor_tflops": flop / (ms * 1e-3) / 1e12,
            "torch_cuda_fp32_ms_per_step": ref_ms, "speedup_vs_torch_cuda_fp32": ref_ms / ms}


def parity_leg(dev, pose, H, W, focal, base_z, ids, precisions):
    """The ray-by-ray parity theorem (tests/parity_tools.py) on 8,192 rays of the timed view, plus PSNR figures."""
    from nerf_b200 import ops
    from oracle import nerf_oracle as O
    from tests.parity_tools import render_parity_report
    sp, sn = synthetic_state_dicts()
    sp, sn = O.params_to(sp, dev), O.params_to(sn, dev)
    R = 8192
    n = H * W
    rays = ops.generate_rays(pose, H, W, focal, focal)[n // 2: n // 2 + R].contiguous()
    g = torch.Generator().manual_seed(7)
    jit, u = torch.rand(R, N_COARSE, generator=g).to(dev), torch.rand(R, N_FINE + 1, generator=g).to(dev)
    ref = O.render_rays(sp, sn, rays, base_z, jit, u, NEAR, FAR, N_FINE, white_bkg=True)
    gt = ref["rgb"] + 0.0316 * torch.randn(ref["rgb"].shape, generator=torch.Generator().manual_seed(1)).to(dev)   # a ground truth 30 dB away

    def psnr(a):
        return -10.0 * math.log10(float(((a - gt) ** 2).mean()))
    par = {"rays": R}
    for mode in precisions:
        got = ops.render_rays(rays, base_z, NEAR, FAR, N_FINE, white_bkg=True, precision=mode, jitter=jit, u=u, debug=True, **ids)
        err = (got["rgb"] - ref["rgb"]).abs()
        mse = float((err ** 2).mean())
        par[mode] = {"max_abs_rgb_err": float(err.max()), "frac_rays_over_1e-4": float((err.amax(-1) > 1e-4).float().mean()),
                     "psnr_vs_reference_db": (99.0 if mse == 0 else -10.0 * math.log10(mse)),
                     "psnr_delta_db_at_30dB_gt": psnr(got["rgb"]) - psnr(ref["rgb"])}
        if mode in ("fp16x3", "bf16x3", "fp32"):
            par[mode]["theorem"] = render_parity_report(O, sp, sn, rays, base_z, jit, u, NEAR, FAR, got, dep_tol=1e-4 if mode == "fp32" else 1.5e-4)
    return par


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--precision", default="fp16x3", choices=["fp32", "fp16x3", "bf16x3", "fp16", "bf16"])
    ap.add_argument("--size", type=int, default=0, help="image side; default 400 (N = 1, weak) / 800 (strong)")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary precision modes / parity / baselines")
    ap.add_argument("--train-ddp", action="store_true",
                    help="N > 1: also time the data-parallel training step (ddp_train.py: per-rank ray batches, one flat gradient all-reduce)")
    ap.add_argument("--scaling", default="auto", choices=["auto", "weak", "strong"],
                    help="auto: strong for N > 1 (ONE 800x800 image sharded across the GPUs, BASELINE configs[4]); weak: one view per GPU")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="strong scaling: rgb rows stored into every GPU's image from the compositing epilogue (peer) | NCCL all_gather")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch.distributed as dist
    import nerf_b200
    from nerf_b200 import _lib, ops, sharding
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W_ = max(args.warmup, 3)
    K = args.steps
    strong = (args.scaling == "strong") or (args.scaling == "auto" and world > 1)
    H = Wd = args.size if args.size else (800 if strong else 400)
    n_rays = H * Wd

    # ---- model + inputs (random-init weights of the reference architecture, synthetic orbit poses) ----
    sd_prop, sd_nerf = synthetic_state_dicts()
    prop = nerf_b200.ProposalNetwork(10, 256)
    net = nerf_b200.MipNeRF(10, 4, 256)
    prop.load_state_dict(sd_prop)
    net.load_state_dict(sd_nerf)
    prop, net = prop.to(dev), net.to(dev)
    with torch.no_grad():
        ids = dict(prop_net_id=prop._nb2_sync(), nerf_net_id=net._nb2_sync())
    base_z = torch.linspace(NEAR, FAR, N_COARSE, device=dev)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)    # > 126 MB L2
    lib, h = _lib.load(), _lib.handle(dev)

    class Scene:
        """One workload: geometry of the view, this rank's shard, pre-allocated buffers (nothing allocates in a step)."""

        def __init__(self, H, W, strong, gather):
            self.H, self.W, self.strong, self.gather = H, W, strong, gather
            self.n = H * W
            self.start, self.count = sharding.shard_range(self.n, rank, world) if strong else (0, self.n)
            theta = 30.0 if strong else -180.0 + 360.0 * rank / max(world, 1) + 30.0
            self.pose_host = nerf_b200.pose_spherical(theta, -30.0, 4.0)[:3, :].contiguous().pin_memory()
            self.pose = self.pose_host.to(dev)
            self.focal = float(nerf_b200.fov2Focal(FOV, (H, W))[0])
            self.rays = torch.empty((self.count, 6), dtype=torch.float32, device=dev)
            self.peer = self.gbuf = self.weak_out = None
            self.out = None
            self.ws = None
            if strong and world > 1 and gather == "peer":
                self.peer = sharding.PeerImage(self.n, 3, dev)
                self.out = {"rgb": self.peer.local_rows(self.start, self.count),
                            "depth": torch.empty((self.count,), dtype=torch.float32, device=dev),
                            "acc": torch.empty((self.count,), dtype=torch.float32, device=dev)}
            elif strong and world > 1:
                self.gbuf = sharding.GatherBuffers(self.n, 3, world, dev)
            elif world > 1:
                self.weak_out = torch.empty((world * self.n, 3), dtype=torch.float32, device=dev)

        def step(self, precision, seed, pose_t=None):
            """ray generation + the three launches of nb2_render_rays (+ the gather).  Returns the full image rows on
            every rank in strong mode, this rank's own view otherwise."""
            pose_t = self.pose if pose_t is None else pose_t
            ops.generate_rays(pose_t, self.H, self.W, self.focal, self.focal, pix_offset=self.start, n_rays=self.count, out=self.rays)
            # device RNG is keyed on the GLOBAL ray id (ray_offset), so the image does not depend on the world size
            out = ops.render_rays(self.rays, base_z, NEAR, FAR, N_FINE, white_bkg=True, precision=precision, seed=seed,
                                  ray_offset=self.start, workspace=self.ws, out=self.out,
                                  peer_rgb=self.peer.peer_ptrs if self.peer is not None else (), **ids)
            self.out, self.ws = out, out["_workspace"]
            if self.peer is not None:
                self.peer.fence()                    # every rank's rows have landed in every image
                return self.peer.image
            if self.gbuf is not None:
                return sharding.gather_rows(out["rgb"], self.n, buffers=self.gbuf)     # the one exchange: final tile gather
            if self.weak_out is not None:
                dist.all_gather_into_tensor(self.weak_out, out["rgb"])
            return out["rgb"]

        def close(self):
            if self.peer is not None:
                self.peer.close()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(scene, precision, steps, kernel_events=False, fn=None):
        fn = fn or (lambda i: scene.step(precision, i))
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        kev = None
        if kernel_events:
            import ctypes
            kev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
        for i in range(W_):
            fn(i)
        barrier()
        n0 = lib.nb2_launch_count(h)
        for i in range(steps):
            flush.zero_()                                      # L2 flush between timed iterations (not timed)
            if kev is not None:
                for e in kev[i]:
                    e.record()                                 # materialise the cudaEvent_t handles
                arr = (ctypes.c_void_p * 4)(*[e.cuda_event for e in kev[i]])
                _lib.check(lib.nb2_set_profile_events(h, arr))
            evs[i][0].record()
            fn(1000 + i)
            evs[i][1].record()
        barrier()
        if kev is not None:
            _lib.check(lib.nb2_set_profile_events(h, None))
        launches = lib.nb2_launch_count(h) - n0
        ms = reduce_max(sum(a.elapsed_time(b) for a, b in evs) / steps)
        kms = [sum(k[j].elapsed_time(k[j + 1]) for k in kev) / steps for j in range(3)] if kev is not None else None
        return ms, launches, kms

    with torch.no_grad():
        scene = Scene(H, Wd, strong, args.gather)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ms, launches, kms = timed(scene, args.precision, K, kernel_events=True)
        clocks = sampler.stop() if rank == 0 else None
        # sanity on what the timed loop produced (not timed): every pixel of the last step finite and inside [0, 1] + eps
        last = scene.step(args.precision, 1000 + K - 1)
        torch.cuda.synchronize()
        sane = torch.tensor([float(torch.isfinite(last).all() and float(last.min()) > -1e-3 and float(last.max()) < 1.0 + 1e-3)], device=dev)
        if world > 1:
            dist.all_reduce(sane, op=dist.ReduceOp.MIN)
        if float(sane.item()) != 1.0:
            raise SystemExit("bench.py: the rendered image contains non-finite or out-of-range pixels")
        total_rays = n_rays if strong else world * n_rays
        value = total_rays / (ms / 1e3)

        # ---- e2e: pose from pinned host memory, FULL image back to pinned host memory, every step ----
        img_host = torch.empty((n_rays, 3) if strong else (3, H, Wd), dtype=torch.float32).pin_memory()

        def e2e_step(i):
            p_dev = scene.pose_host.to(dev, non_blocking=True)
            if strong:
                # the sharded render has no single-call API in the reference's surface: pose in, shard, gather, image out
                img = scene.step(args.precision, i, p_dev)
                if rank == 0:
                    img_host.copy_(img, non_blocking=True)
                return
            res = nerf_b200.render_image(net, prop, p_dev, (H, Wd), scene.focal, NEAR, FAR, N_FINE, white_bkg=True, precision=args.precision, seed=i)
            img_host.copy_(res["rgb"], non_blocking=True)
        e_ms, _, _ = timed(scene, args.precision, K, fn=e2e_step)
        e2e = {"value": total_rays / (e_ms / 1e3), "unit": "rays/s", "h2d_bytes_per_step": scene.pose_host.numel() * 4,
               "d2h_bytes_per_step": img_host.numel() * 4, "ms_per_step": e_ms,
               "api": "pose (pinned host) -> shard render -> fused gather -> full image to pinned host on rank 0" if strong
                      else "nerf_b200.render_image (pose from pinned host memory, image to pinned host memory)"}

        extras, multi = {}, {}
        if strong and world > 1 and not args.no_extras:
            # (a) the gathered image must not depend on the world size: rank 0 renders all rays itself with the same seed
            img = scene.step(args.precision, 4242).clone()
            torch.cuda.synchronize()
            if rank == 0:
                full = ops.render_rays(ops.generate_rays(scene.pose, H, Wd, scene.focal, scene.focal), base_z, NEAR, FAR, N_FINE,
                                       white_bkg=True, precision=args.precision, seed=4242, **ids)
                multi["shard_invariance_max_abs_diff"] = float((img - full["rgb"]).abs().max())
            barrier()
            # (b) the other gather implementation on the same workload
            other = "nccl" if args.gather == "peer" else "peer"
            sc2 = Scene(H, Wd, True, other)
            o_ms, _, _ = timed(sc2, args.precision, max(5, K // 2))
            sc2.close()
            multi[f"gather_{other}"] = {"ms_per_step": o_ms, "rays_per_s": n_rays / (o_ms / 1e3)}
            multi[f"gather_{args.gather}"] = {"ms_per_step": ms, "rays_per_s": value}
            # (c) weak scaling (round 1's multi-GPU line): one 400x400 view per GPU, tiles all-gathered
            sc3 = Scene(400, 400, False, "nccl")
            w_ms, _, _ = timed(sc3, args.precision, max(5, K // 2))
            multi["weak_scaling"] = {"workload": "one 400x400 view per GPU, all_gather of the tiles", "ms_per_step": w_ms,
                                     "rays_per_s": world * 160000 / (w_ms / 1e3)}
        if not args.no_extras:
            for mode in ("bf16", "fp16"):
                if mode == args.precision:
                    continue
                m_ms, _, m_k = timed(scene, mode, max(5, K // 2), kernel_events=True)
                extras[mode] = {"rays_per_s": total_rays / (m_ms / 1e3), "ms_per_step": m_ms, "fine_kernel_ms": m_k[2],
                                "fine_kernel_tflops": FLOP_NERF_PER_RAY * scene.count / (m_k[2] * 1e-3) / 1e12}

    train_ddp = None
    if world > 1 and args.train_ddp:
        with torch.enable_grad():
            train_ddp = train_step_leg(dev, ddp=(rank, world))
    if rank != 0:
        scene.close()
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    fine_ms = kms[2]
    achieved = FLOP_NERF_PER_RAY * scene.count / (fine_ms * 1e-3) / 1e12     # rank 0's launch
    passes = 3 if args.precision in ("fp16x3", "bf16x3") else 1
    traffic = NCU_FINE_TRAFFIC_BYTES_PER_RAY.get(args.precision)
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                "traffic": traffic * scene.count if traffic else None,
                "traffic_unit": "bytes per launch (ncu dram read+write per ray x rays of the launch; algorithmic 556 B/ray)",
                "kernel": FINE_KERNEL[args.precision], "kernel_ms": fine_ms,
                "algorithmic_flop_per_launch": FLOP_NERF_PER_RAY * scene.count, "peak_source": peaks["src"], "mma_passes_per_product": passes,
                "issued_tensor_tflops": achieved * passes * 528384.0 / 527872.0,
                "frac_issued": achieved * passes * 528384.0 / 527872.0 / peaks["tflops"],
                "frac_of_precision_ceiling": achieved * passes / peaks["tflops"],
                "step_share": {"proposal_kernel_ms": kms[0], "resample_kernel_ms": kms[1], "fine_kernel_ms": kms[2]}}
    for mode, m in extras.items():
        m["fine_kernel_frac_of_peak"] = m["fine_kernel_tflops"] / peaks["tflops"]

    if strong:
        workload = (f"ONE Lego-shaped {H}x{Wd} orbit view per step, its {n_rays} rays sharded contiguously across {world} GPU(s) and gathered into "
                    "the full image on every rank (BASELINE configs[4]), ")
        par = f"ray-sharded x{world}; gather = " + ("rgb rows stored into every GPU's image by the compositing epilogue over NVLink (CUDA IPC), one 4-byte all_reduce as fence"
                                                      if args.gather == "peer" and world > 1 else "all_gather_into_tensor of the rgb rows")
    else:
        workload = f"Lego-shaped {H}x{Wd} orbit view per GPU (BASELINE configs[1]), "
        par = f"independent views x{world}" + (", all_gather of the rgb tiles" if world > 1 else "")
    line = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": K, "warmup": W_, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": DTYPE[args.precision], "data": "synthetic",
        "config": {"workload": workload + "64 coarse + 128 fine samples, proposal 4x256 + NeRF 8x256",
                   "rays_per_gpu_per_step": scene.count, "precision": args.precision, "l2": "flushed between timed steps (512 MB memset)",
                   "rng": "device Philox keyed on global ray id", "parallelism": par, **multi},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "other_precisions": extras,
    }

    if train_ddp is not None:
        line["train_step_ddp"] = train_ddp
    if not args.no_extras and world == 1:
        def leg(name, fn, *a, grad=False):       # an extra leg that fails must not take the headline line with it
            try:
                with torch.set_grad_enabled(grad):
                    return fn(*a)
            except Exception as e:               # noqa: BLE001 -- reported in the line, rank 0 stderr has the traceback
                import traceback
                traceback.print_exc()
                return {"error": f"{name}: {type(e).__name__}: {e}"}
        line["parity_vs_oracle"] = leg("parity", parity_leg, dev, scene.pose, H, Wd, scene.focal, base_z, ids,
                                       list(dict.fromkeys([args.precision, "bf16", "fp16"])))
        line["torch_cuda_baseline"] = leg("torch_cuda_baseline", torch_cuda_baseline, dev, 400, 400, grad=True)
        line["train_step"] = leg("train_step", train_step_leg, dev, grad=True)
        line["other_configs"] = {"config3_ipe": leg("config3", config3_leg, dev, prop, net, ids),
                                 "config4_refnerf": leg("config4", config4_leg, dev, prop)}
        cores = os.cpu_count() or 1
        cpu_v, cpu_t = cpu_reference_render(2, 1, cores)
        line["cpu_baseline"] = {"value": cpu_v, "unit": "rays/s", "cores": cores, "kind": "port",
                                "sample": f"2 x a full {CPU_SAMPLE_HW}x{CPU_SAMPLE_HW} view of the same scene (four 50x50 reference tiles, {cpu_t:.2f} s each), PyTorch CPU fp32"}
    print(json.dumps(line))
    scene.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
