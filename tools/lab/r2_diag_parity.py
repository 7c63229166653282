"""Lab: per-ray arrays of the parity report at 400x400 (fp16x3 and fp32), saved for offline analysis."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "."))
import nerf_b200
from nerf_b200 import ops
from oracle import nerf_oracle as O
from tests.parity_tools import render_parity_report
DEV = "cuda"
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 400
pose = nerf_b200.pose_spherical(30.0, -30.0, 4.0)[:3, :].to(DEV)
focal = nerf_b200.fov2Focal(0.6911112070083618, (H, W))[0]
g = torch.Generator(device="cpu").manual_seed(1234)
jitter = torch.rand(H * W, 64, generator=g).to(DEV)
u = torch.rand(H * W, 129, generator=g).to(DEV)
prop = nerf_b200.ProposalNetwork(10, 256); net = nerf_b200.MipNeRF(10, 4, 256)
prop.load_state_dict(O.make_params("proposal", 1, "smooth")); net.load_state_dict(O.make_params("nerf", 2, "smooth"))
prop, net = prop.to(DEV), net.to(DEV)
sp, sn = O.params_to(O.make_params("proposal", 1, "smooth"), DEV), O.params_to(O.make_params("nerf", 2, "smooth"), DEV)
rays = ops.generate_rays(pose, H, W, focal, focal)
base = torch.linspace(2.0, 6.0, 64, device=DEV)
with torch.no_grad():
    ids = dict(nerf_net_id=net._nb2_sync(), prop_net_id=prop._nb2_sync())
    for prec in ("fp16x3", "fp32"):
        eng = ops.render_rays(rays, base, 2.0, 6.0, 128, white_bkg=True, precision=prec, jitter=jitter, u=u, debug=True, **ids)
        arr = {}
        rep = render_parity_report(O, sp, sn, O.generate_rays(pose, H, W, focal), base, jitter, u, 2.0, 6.0, eng, chunk=8000, arrays=arr)
        print(prec, rep)
        np.savez_compressed(f"gpurun_out/r2_diag_{prec}_{H}.npz", **{k: v.numpy() for k, v in arr.items() if v.numel() == H * W})
