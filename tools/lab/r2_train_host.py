"""Lab: where the host time of a 1024-ray training step goes (cProfile)."""
import sys, os, cProfile, pstats, io, time
import torch
import torch.nn.functional as F
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "."))
import nerf_b200
from nerf_b200 import NeRF, ProposalNetwork, getBounds, inverseSample, maxBlurFilter, synthetic
R = 1024
dev = "cuda"
prop_net, mip_net = nerf_b200.ProposalNetwork(10, 256), nerf_b200.MipNeRF(10, 4, 256)
prop_net.load_state_dict(synthetic.make_params("proposal", 1, "smooth")); mip_net.load_state_dict(synthetic.make_params("nerf", 2, "smooth"))
prop_net, mip_net = prop_net.to(dev), mip_net.to(dev)
opt = torch.optim.Adam(list(mip_net.parameters()) + list(prop_net.parameters()), lr=1.5e-4)
Hh = Ww = 400
rgbs = torch.rand(Hh * Ww, 3, device=dev)
rows, cols = torch.meshgrid(torch.arange(Hh), torch.arange(Ww), indexing="ij")
coords = torch.stack((cols - Ww // 2, Hh // 2 - rows), dim=-1).reshape(-1, 2).to(dev)
cam_tf = nerf_b200.pose_spherical(30.0, -30.0, 4.0)[:3, :].contiguous().to(dev)
focal = nerf_b200.fov2Focal(0.6911112070083618, (Hh, Ww))
pl, sl = nerf_b200.ProposalLoss(), nerf_b200.SoftL1Loss()
def step():
    cs, cl, rt, cr = nerf_b200.validSampler(rgbs, coords, cam_tf, R, 64, focal, 2.0, 6.0, True)
    density = F.softplus(prop_net.forward(cs))
    pw = maxBlurFilter(ProposalNetwork.get_weights(density, cl, cr[:, 3:]), 0.01)
    fl, below = inverseSample(pw, cl, 129, sort=True)
    fl = fl[..., :-1]
    rgbo = mip_net.forward(NeRF.length2pts(cr, fl))
    rendered, weights, _ = NeRF.render(rgbo, fl, cr[:, 3:])
    wb = getBounds(pw, below)
    opt.zero_grad()
    loss = pl(wb, weights.detach()) + sl(rendered, rt)
    loss.backward()
    opt.step()
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): step()
torch.cuda.synchronize()
print("ms per step", (time.perf_counter() - t0) / 20 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
