"""Lab: per-parameter gradient errors of the layer-wise engine vs fp64 autograd on the oracle."""
import sys, os
import torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "."))
import nerf_b200
from oracle import nerf_oracle as O
DEV = "cuda"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
style = sys.argv[2] if len(sys.argv) > 2 else "smooth"
sp, sn = O.make_params("proposal", 1, style), O.make_params("nerf", 2, style)
pts = torch.cat((O.det_uniform((n, 3), 9, -2.0, 2.0), O.det_uniform((n, 3), 10, -1.0, 1.0)), -1).to(DEV)
g_rgbo = O.det_uniform((n, 4), 11, -1.0, 1.0).to(DEV)
g_sig = O.det_uniform((n,), 12, -1.0, 1.0).to(DEV)
sp64 = {k: v.to(DEV).double().requires_grad_(True) for k, v in sp.items()}
sn64 = {k: v.to(DEV).double().requires_grad_(True) for k, v in sn.items()}
(O.nerf_forward(sn64, pts.double()) * g_rgbo.double()).sum().backward()
(O.proposal_forward(sp64, pts[:, :3].double()) * g_sig.double()).sum().backward()
def ne(a, b):
    return float((a.double() - b.double()).norm()) / max(float(b.double().norm()), 1e-30)
for prec in ("bf16x3", "bf16"):
    prop = nerf_b200.ProposalNetwork(10, 256); net = nerf_b200.MipNeRF(10, 4, 256)
    prop.load_state_dict(sp); net.load_state_dict(sn)
    prop, net = prop.to(DEV), net.to(DEV)
    prop.train_precision = net.train_precision = prec
    out = net.forward(pts[None]); (out[0] * g_rgbo).sum().backward()
    d = prop.forward(pts[None, :, :3].contiguous()); (d[0] * g_sig).sum().backward()
    print("==", prec, style, n)
    for m, r in ((net, sn64), (prop, sp64)):
        for k, p in m.named_parameters():
            print(f"  {k:24s} norm-err {ne(p.grad, r[k].grad):.3e}   |ref| {float(r[k].grad.norm()):.3e}")
