"""Lab: the standalone HBM-bound ops at BASELINE config-2 sizes (160,000 rays), one launch each after an L2 flush; run under
`ncu --metrics gpu__time_duration.sum` and post-process with tools/lab/r2_hbm_table.py (kernel name -> algorithmic bytes)."""
import sys, os, json, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "."))
import nerf_b200
from nerf_b200 import ops
DEV = "cuda"
R = 160000
g = torch.Generator().manual_seed(3)
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=DEV)
pose = nerf_b200.pose_spherical(30.0, -30.0, 4.0)[:3, :].to(DEV)
focal = float(nerf_b200.fov2Focal(0.6911112070083618, (400, 400))[0])
rays = ops.generate_rays(pose, 400, 400, focal, focal)
base_z = torch.linspace(2.0, 6.0, 64, device=DEV)
jitter = torch.rand(R, 64, generator=g).to(DEV)
u = torch.rand(R, 129, generator=g).to(DEV)
z, pts = ops.sample_coarse(rays, base_z, 4.0 / 128, jitter=jitter)
sigma = (torch.rand(R, 64, generator=g) * 30.0).to(DEV)
dirs = rays[:, 3:].contiguous()
w = ops.weights_from_sigma(sigma, z, dirs)
zf = ops.resample(sigma, z, rays, 129, u=u)
rgbo = torch.rand(R, 128, 4, generator=g).to(DEV)
x3 = pts.view(-1, 3)[: R * 8].contiguous()
cases = [
    ("generate_rays_kernel", lambda: ops.generate_rays(pose, 400, 400, focal, focal), 160000 * 24),
    ("sample_coarse", lambda: ops.sample_coarse(rays, base_z, 4.0 / 128, jitter=jitter), R * (24 + 256 + 256 + 768)),
    ("posenc", lambda: ops.posenc(x3, 10), x3.shape[0] * (12 + 240)),
    ("ipe_kernel", lambda: ops.ipe(z, rays, 10, 0.01), R * 24 + R * 64 * 4 + R * 63 * (240 + 16)),
    ("weights", lambda: ops.weights_from_sigma(sigma, z, dirs), R * (256 + 256 + 12 + 256)),
    ("max_blur", lambda: ops.max_blur(w, 0.01), R * 512),
    ("resample_kernel", lambda: ops.resample(sigma, z, rays, 129, u=u), R * (256 + 256 + 24 + 516 + 512)),
    ("length2pts", lambda: ops.length2pts(rays, zf), R * (24 + 512 + 128 * 24)),
    ("composite", lambda: ops.composite(rgbo, zf, dirs, white_bkg=True, near_far=(2.0, 6.0), want_weights=False), R * (128 * 20 + 12 + 16)),
]
json.dump({k: b for k, _, b in cases}, open("gpurun_out/r2_hbm_bytes.json", "w"))
for name, fn, _ in cases:
    fn(); fn()
torch.cuda.synchronize()
for rep in range(3):
    for name, fn, _ in cases:
        flush.zero_()
        fn()
torch.cuda.synchronize()
print("done")
