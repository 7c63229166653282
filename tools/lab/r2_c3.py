import sys, os, json, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "."))
import bench, nerf_b200
dev = torch.device("cuda", 0)
sd_prop, sd_nerf = bench.synthetic_state_dicts()
prop = nerf_b200.ProposalNetwork(10, 256); net = nerf_b200.MipNeRF(10, 4, 256)
prop.load_state_dict(sd_prop); net.load_state_dict(sd_nerf)
prop, net = prop.to(dev), net.to(dev)
with torch.no_grad():
    ids = dict(prop_net_id=prop._nb2_sync(), nerf_net_id=net._nb2_sync())
    print(json.dumps(bench.config3_leg(dev, prop, net, ids), indent=1))
