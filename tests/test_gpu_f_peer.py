"""GPU x2 (skipped on a single-GPU box): the fused multi-GPU gather.  Two ranks (one process per GPU, NCCL) render the
two halves of one image; the compositing epilogue stores every rgb row into BOTH GPUs' image buffers over NVLink
(nerf_b200.sharding.PeerImage, nb2_render_params.peer_rgb).  Both ranks must end up with the image a single GPU renders."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import nerf_b200
    from nerf_b200 import ops, sharding, synthetic
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    prop = nerf_b200.ProposalNetwork(10, 256)
    net = nerf_b200.MipNeRF(10, 4, 256)
    prop.load_state_dict(synthetic.make_params("proposal", 1, "smooth"))
    net.load_state_dict(synthetic.make_params("nerf", 2, "smooth"))
    prop, net = prop.to(dev), net.to(dev)
    H = W = 90                                       # 8100 rays: ragged against the shard alignment
    pose = nerf_b200.pose_spherical(30.0, -30.0, 4.0)[:3, :].to(dev)
    focal = float(nerf_b200.fov2Focal(0.6911112070083618, (H, W))[0])
    base = torch.linspace(2.0, 6.0, 64, device=dev)
    results = {}
    with torch.no_grad():
        ids = dict(prop_net_id=prop._nb2_sync(), nerf_net_id=net._nb2_sync())
        n = H * W
        start, count = sharding.shard_range(n, rank, world)
        peer = sharding.PeerImage(n, 3, dev)
        for precision in ("fp16x3", "bf16", "fp32"):
            peer.image.zero_()
            dist.barrier()
            rays = ops.generate_rays(pose, H, W, focal, focal, pix_offset=start, n_rays=count)
            out = {"rgb": peer.local_rows(start, count), "depth": torch.empty(count, device=dev), "acc": torch.empty(count, device=dev)}
            ops.render_rays(rays, base, 2.0, 6.0, 128, white_bkg=True, precision=precision, seed=5, ray_offset=start, out=out,
                            peer_rgb=peer.peer_ptrs, **ids)
            peer.fence()
            torch.cuda.synchronize()
            full = ops.render_rays(ops.generate_rays(pose, H, W, focal, focal), base, 2.0, 6.0, 128, white_bkg=True, precision=precision,
                                   seed=5, **ids)["rgb"]
            results[precision] = bool(torch.equal(peer.image, full))
            # the NCCL gather of the same rows gives the same image
            g = sharding.gather_rows(out["rgb"].clone(), n)
            results[precision + "_nccl"] = bool(torch.equal(g, full))
        peer.close()
    q.put((rank, results))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one node")
def test_peer_store_gather_equals_single_gpu_image():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    for rank, res in got:
        assert all(res.values()), (rank, res)
