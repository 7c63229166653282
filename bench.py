#!/usr/bin/env python
"""bench.py — rays/s of the render hot path (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            the engine (libnerfb200.so)
  python bench.py --impl reference ...                     the reference algorithm on the host CPU cores
                                                           (the oracle port; /root/reference does not travel)
Workload (BASELINE.json configs[1]): Lego-shaped 400x400 orbit view, 64 coarse + 128 fine samples per ray,
vanilla NeRF 8x256 + 4x256 proposal MLP, fp32-faithful arithmetic, synthetic poses and random-init weights
(band-limited 'smooth' field, oracle/nerf_oracle.py:make_params).  A step = one pass of the hot path over one
400x400 ray batch per GPU (weak scaling: every rank renders its own view; the rendered tiles are all-gathered).

One JSON line on stdout (rank 0).  `value` = rays/s with inputs resident in HBM (CUDA events, max over ranks,
L2 flushed between timed steps); `e2e` = the same through the public API with the pose in pinned host memory
and the image copied back to the host inside the timed region; `roofline` = the fused encode+MLP+composite
kernel against the measured bf16 tensor peak; `cpu_baseline` = the oracle on this box's host cores.
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
# dram__bytes_read.sum + dram__bytes_write.sum of ONE fine-kernel launch over 160,000 rays, from the ncu --set full
# captures summarised in profiles/r01_ncu_{fp16x3_tmema,fp16_pp}.txt (algorithmic: 556 B/ray = 89 MB; the
# 2.4 MB of packed weights stay L2-resident)
NCU_FINE_TRAFFIC_BYTES_160K = {"fp16x3": 93.1e6, "bf16x3": 93.1e6, "fp16": 91.9e6, "bf16": 91.9e6}
FINE_KERNEL = {"fp16x3": "mlp_tc4_kernel (fine: encode + 8x256 MLP + composite; CTA-pair tcgen05, activations in TMEM)",
               "bf16x3": "mlp_tc4_kernel (fine: encode + 8x256 MLP + composite; CTA-pair tcgen05, activations in TMEM)",
               "fp16": "mlp_tc2_kernel (fine: encode + 8x256 MLP + composite; CTA-pair tcgen05, ping-pong tiles)",
               "bf16": "mlp_tc2_kernel (fine: encode + 8x256 MLP + composite; CTA-pair tcgen05, ping-pong tiles)",
               "fp32": "mlp_simt_kernel (fine: encode + 8x256 MLP on CUDA cores) + composite_kernel"}


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


def oracle_setup(device):
    from oracle import nerf_oracle as O
    sp, sn = O.make_params("proposal", 1, "smooth"), O.make_params("nerf", 2, "smooth")
    return O, O.params_to(sp, device), O.params_to(sn, device)


def cpu_oracle_rays_per_s(n_rays, reps, threads):
    """The reference algorithm (oracle port, PyTorch fp32 CPU ops = the reference's own arithmetic) on the host."""
    import nerf_b200
    torch.set_num_threads(threads)
    O, sp, sn = oracle_setup("cpu")
    H = W = 400
    pose = nerf_b200.pose_spherical(30.0, -30.0, 4.0)[:3, :]
    focal = nerf_b200.fov2Focal(FOV, (H, W))[0]
    rays = O.generate_rays(pose, H, W, focal)
    # one reference tile = 50x50 pixels (nerf/procedures.py:21,60-64); take the centre tile(s)
    sel = torch.arange(n_rays) + (H // 2) * W
    rays = rays[sel]
    base_z = torch.linspace(NEAR, FAR, N_COARSE)
    times = []
    with torch.no_grad():
        for i in range(reps + 1):
            g = torch.Generator().manual_seed(i)
            jit, u = torch.rand(n_rays, N_COARSE, generator=g), torch.rand(n_rays, N_FINE + 1, generator=g)   # the CPU draws of the reference
            t0 = time.perf_counter()
            O.render_rays(sp, sn, rays, base_z, jit, u, NEAR, FAR, N_FINE, white_bkg=True, chunk=2500)
            times.append(time.perf_counter() - t0)
    t = sum(times[1:]) / reps
    return n_rays / t, t


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_rays = 2500
    torch.set_num_threads(threads)
    O, sp, sn = oracle_setup("cpu")
    import nerf_b200
    H = W = 400
    pose = nerf_b200.pose_spherical(30.0, -30.0, 4.0)[:3, :]
    focal = nerf_b200.fov2Focal(FOV, (H, W))[0]
    rays = O.generate_rays(pose, H, W, focal)[torch.arange(n_rays) + (H // 2) * W]
    base_z = torch.linspace(NEAR, FAR, N_COARSE)
    times = []
    with torch.no_grad():
        for i in range(args.warmup + args.steps):
            jit, u = torch.rand(n_rays, N_COARSE), torch.rand(n_rays, N_FINE + 1)
            t0 = time.perf_counter()
            O.render_rays(sp, sn, rays, base_z, jit, u, NEAR, FAR, N_FINE, white_bkg=True, chunk=2500)
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    value = n_rays / (ms / 1e3)
    sample = f"{n_rays} rays (one 50x50 reference tile of the 400x400 view) per step, PyTorch {torch.__version__} CPU fp32, {threads} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": "Lego-shaped 400x400 orbit view, 64 coarse + 128 fine, vanilla NeRF (configs[1]); bounded sample", "rays_per_step": n_rays},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--precision", default="fp16x3", choices=["fp32", "fp16x3", "bf16x3", "fp16", "bf16"])
    ap.add_argument("--size", type=int, default=400)
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary precision modes / parity / cpu baseline")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): every GPU renders its own --size x --size view; strong: ONE --size x --size image, its "
                         "rays sharded contiguously across the GPUs and gathered (BASELINE configs[4] at --size 800)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch.distributed as dist
    import nerf_b200
    from nerf_b200 import _lib, ops
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
    H = Wd = args.size
    n_rays = H * Wd
    strong = args.scaling == "strong"
    from nerf_b200 import sharding
    # strong scaling: this rank's contiguous slice of the one image
    r_start, r_count = sharding.shard_range(n_rays, rank, world) if strong else (0, n_rays)

    # ---- model + inputs (random-init weights of the reference architecture, synthetic orbit poses) ----
    from oracle import nerf_oracle as O   # weight generator only here; the checker use is in parity_check()
    prop = nerf_b200.ProposalNetwork(10, 256)
    net = nerf_b200.MipNeRF(10, 4, 256)
    prop.load_state_dict(O.make_params("proposal", 1, "smooth"))
    net.load_state_dict(O.make_params("nerf", 2, "smooth"))
    prop, net = prop.to(dev), net.to(dev)
    with torch.no_grad():
        prop._nb2_sync(); net._nb2_sync()
    theta = 30.0 if strong else -180.0 + 360.0 * rank / max(world, 1) + 30.0
    pose_host = nerf_b200.pose_spherical(theta, -30.0, 4.0)[:3, :].contiguous().pin_memory()
    pose = pose_host.to(dev)
    focal = float(nerf_b200.fov2Focal(FOV, (H, Wd))[0])
    base_z = torch.linspace(NEAR, FAR, N_COARSE, device=dev)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)    # > 126 MB L2
    lib, h = _lib.load(), _lib.handle(dev)
    gathered = torch.empty((world * n_rays, 3), dtype=torch.float32, device=dev) if (world > 1 and not strong) else None
    state = {"ws": None}

    def step(precision, seed, pose_t=None):
        pose_t = pose if pose_t is None else pose_t
        if strong:
            # device RNG is keyed on the GLOBAL ray id (ray_offset), so the image does not depend on the world size
            rays = ops.generate_rays(pose_t, H, Wd, focal, focal, pix_offset=r_start, n_rays=r_count)
            out = ops.render_rays(rays, base_z, NEAR, FAR, N_FINE, white_bkg=True, precision=precision, seed=seed,
                                  ray_offset=r_start, workspace=state["ws"])
            state["ws"] = out["_workspace"]
            out["image_rows"] = sharding.gather_rows(out["rgb"], n_rays)     # the one exchange: final tile gather
            return out
        rays = ops.generate_rays(pose_t, H, Wd, focal, focal)
        out = ops.render_rays(rays, base_z, NEAR, FAR, N_FINE, white_bkg=True, precision=precision, seed=seed, workspace=state["ws"])
        state["ws"] = out["_workspace"]
        if world > 1:
            dist.all_gather_into_tensor(gathered, out["rgb"])     # the one exchange: final tile gather
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(precision, steps, kernel_events=False):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        kev = None
        if kernel_events:
            import ctypes
            kev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
        for i in range(W_):
            step(precision, i)
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
            step(precision, 1000 + i)
            evs[i][1].record()
        barrier()
        if kev is not None:
            _lib.check(lib.nb2_set_profile_events(h, None))
        launches = lib.nb2_launch_count(h) - n0
        ms = sum(a.elapsed_time(b) for a, b in evs) / steps
        kms = None
        if kev is not None:
            kms = [sum(k[j].elapsed_time(k[j + 1]) for k in kev) / steps for j in range(3)]
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches, kms

    with torch.no_grad():
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ms, launches, kms = timed(args.precision, K, kernel_events=True)
        clocks = sampler.stop() if rank == 0 else None
        # sanity on what the timed loop produced (not timed): every pixel of the last step finite and inside [0, 1] + eps
        last = step(args.precision, 1000 + K - 1)["rgb"]
        sane = torch.tensor([float(torch.isfinite(last).all() and float(last.min()) > -1e-3 and float(last.max()) < 1.0 + 1e-3)], device=dev)
        if world > 1:
            dist.all_reduce(sane, op=dist.ReduceOp.MIN)
        if float(sane.item()) != 1.0:
            raise SystemExit("bench.py: the rendered image contains non-finite or out-of-range pixels")
        total_rays = n_rays if strong else world * n_rays
        value = total_rays / (ms / 1e3)

        # ---- e2e: public API, pose from pinned host memory, image back to pinned host memory, every step ----
        img_host = torch.empty((3, H, Wd), dtype=torch.float32).pin_memory()
        e_evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]

        rows_host = torch.empty((n_rays, 3), dtype=torch.float32).pin_memory() if strong else None

        def e2e_step(i):
            p_dev = pose_host.to(dev, non_blocking=True)
            if strong:
                # the sharded render has no single-call public API in the reference's surface: pose in, shard, gather, image out
                res = step(args.precision, i, p_dev)
                if rank == 0:
                    rows_host.copy_(res["image_rows"], non_blocking=True)
                return
            res = nerf_b200.render_image(net, prop, p_dev, (H, Wd), focal, NEAR, FAR, N_FINE, white_bkg=True, precision=args.precision, seed=i)
            img_host.copy_(res["rgb"], non_blocking=True)
        for i in range(W_):
            e2e_step(i)
        barrier()
        for i in range(K):
            flush.zero_()
            e_evs[i][0].record()
            e2e_step(2000 + i)
            e_evs[i][1].record()
        barrier()
        e_ms = sum(a.elapsed_time(b) for a, b in e_evs) / K
        t = torch.tensor([e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_ms = float(t.item())
        e2e = {"value": total_rays / (e_ms / 1e3), "unit": "rays/s", "h2d_bytes_per_step": pose_host.numel() * 4,
               "d2h_bytes_per_step": img_host.numel() * 4, "ms_per_step": e_ms}

        shard_diff = None
        if strong and world > 1 and not args.no_extras:
            # the gathered image must not depend on the world size: rank 0 renders all rays itself with the same seed
            res = step(args.precision, 4242)
            if rank == 0:
                full = ops.render_rays(ops.generate_rays(pose, H, Wd, focal, focal), base_z, NEAR, FAR, N_FINE, white_bkg=True,
                                       precision=args.precision, seed=4242)
                shard_diff = float((res["image_rows"] - full["rgb"]).abs().max())
        extras = {}
        if not args.no_extras:
            for mode in ("bf16", "fp16"):
                if mode == args.precision:
                    continue
                m_ms, _, m_k = timed(mode, max(5, K // 2), kernel_events=True)
                extras[mode] = {"rays_per_s": total_rays / (m_ms / 1e3), "ms_per_step": m_ms,
                                "fine_kernel_ms": m_k[2], "fine_kernel_tflops": FLOP_NERF_PER_RAY * r_count / (m_k[2] * 1e-3) / 1e12}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    fine_ms = kms[2]
    achieved = FLOP_NERF_PER_RAY * r_count / (fine_ms * 1e-3) / 1e12     # rank 0's launch
    passes = 3 if args.precision in ("fp16x3", "bf16x3") else 1
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                "traffic": (NCU_FINE_TRAFFIC_BYTES_160K.get(args.precision) if r_count == 160000 else None),
                "traffic_unit": "bytes per launch (ncu dram read+write; algorithmic 556 B/ray)",
                "kernel": FINE_KERNEL[args.precision], "kernel_ms": fine_ms,
                "algorithmic_flop_per_launch": FLOP_NERF_PER_RAY * r_count, "peak_source": peaks["src"], "mma_passes_per_product": passes,
                "issued_tensor_tflops": achieved * passes * 528384.0 / 527872.0,
                "frac_issued": achieved * passes * 528384.0 / 527872.0 / peaks["tflops"], "step_share": {"proposal_kernel_ms": kms[0], "resample_kernel_ms": kms[1], "fine_kernel_ms": kms[2]}}
    for mode, m in extras.items():
        m["fine_kernel_frac_of_peak"] = m["fine_kernel_tflops"] / peaks["tflops"]

    line = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": K, "warmup": W_, "ms_per_step": ms,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": {"fp16x3": "f32-faithful: fp16 hi+lo split operands, 3 tcgen05 MMAs per product, f32 accumulate", "bf16x3": "bf16 hi+lo split, f32 accumulate",
                  "fp32": "f32 (CUDA cores)", "bf16": "bf16 operands, f32 accumulate", "fp16": "f16 operands, f32 accumulate"}[args.precision],
        "data": "synthetic",
        "config": {"workload": (f"ONE Lego-shaped {H}x{Wd} orbit view, rays sharded across the GPUs (BASELINE configs[4] at 800x800), " if strong
                                else f"Lego-shaped {H}x{Wd} orbit view per GPU, ") + "64 coarse + 128 fine samples, proposal 4x256 + NeRF 8x256 (BASELINE configs[1])",
                   "rays_per_gpu_per_step": r_count, "precision": args.precision, "l2": "flushed between timed steps (512 MB memset)",
                   "rng": "device Philox keyed on global ray id", "parallelism": f"ray-sharded x{world}, all_gather of rgb tiles"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "other_precisions": extras,
    }
    if shard_diff is not None:
        line["config"]["shard_invariance_max_abs_diff"] = shard_diff   # gathered image vs the same image rendered on one GPU

    if not args.no_extras and world == 1:
        # parity spot check against the oracle running the reference algorithm in PyTorch fp32 on the same GPU
        O2, sp, sn = oracle_setup(dev)
        with torch.no_grad():
            R = 8192
            rays = ops.generate_rays(pose, H, Wd, focal, focal)[n_rays // 2: n_rays // 2 + R].contiguous()
            g = torch.Generator().manual_seed(7)
            jit, u = torch.rand(R, N_COARSE, generator=g).to(dev), torch.rand(R, N_FINE + 1, generator=g).to(dev)
            ref = O2.render_rays(sp, sn, rays, base_z, jit, u, NEAR, FAR, N_FINE, white_bkg=True)
            par = {}
            for mode in dict.fromkeys([args.precision, "bf16", "fp16"]):
                got = ops.render_rays(rays, base_z, NEAR, FAR, N_FINE, white_bkg=True, precision=mode, jitter=jit, u=u)
                err = (got["rgb"] - ref["rgb"]).abs()
                mse = float((err ** 2).mean())
                # PSNR delta against a synthetic ground truth 30 dB away from the reference render
                gt = (ref["rgb"] + 0.0316 * torch.randn(ref["rgb"].shape, generator=torch.Generator().manual_seed(1)).to(dev))
                psnr = lambda a: -10.0 * math.log10(float(((a - gt) ** 2).mean()))
                par[mode] = {"max_abs_rgb_err": float(err.max()), "frac_rays_over_1e-4": float((err.amax(-1) > 1e-4).float().mean()),
                             "psnr_vs_reference_db": (99.0 if mse == 0 else -10.0 * math.log10(mse)),
                             "psnr_delta_db_at_30dB_gt": psnr(got["rgb"]) - psnr(ref["rgb"])}
            line["parity_vs_oracle"] = {"rays": R, **par}
        cores = os.cpu_count() or 1
        cpu_v, cpu_t = cpu_oracle_rays_per_s(2500, 3, cores)
        line["cpu_baseline"] = {"value": cpu_v, "unit": "rays/s", "cores": cores, "kind": "port",
                                "sample": f"3 x 2500 rays (one 50x50 reference tile of the same view), {cpu_t:.2f} s each, PyTorch CPU fp32"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
