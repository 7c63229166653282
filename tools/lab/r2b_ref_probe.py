"""Lab: the Ref-NeRF branch of render_image at the config-4 batch (512 rays) and at 10,000 rays, for an ncu launch list."""
import sys, os, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "."))
import nerf_b200
from nerf_b200 import synthetic
dev = "cuda"
torch.manual_seed(0)
rn = nerf_b200.RefNeRF(10, 4)
rn.load_state_dict(synthetic.det_state_dict(rn, 7, gain=1.0))
rn = rn.to(dev).eval()
prop = nerf_b200.ProposalNetwork(10, 256)
prop.load_state_dict({k: v for k, v in synthetic.det_state_dict(prop, 1, gain=1.0).items()})
prop = prop.to(dev).eval()
pose = nerf_b200.pose_spherical(30.0, -30.0, 4.0)[:3, :].to(dev)
shape = (16, 32) if len(sys.argv) < 2 else (int(sys.argv[1]), int(sys.argv[2]))
focal = float(nerf_b200.fov2Focal(0.6911112070083618, (shape[1], shape[1]))[0])
for _ in range(4):
    nerf_b200.render_image(rn, prop, pose, shape, focal, 2.0, 6.0, 128, white_bkg=True, render_depth=True, render_normal=True, seed=3)
torch.cuda.synchronize()
print("done")
