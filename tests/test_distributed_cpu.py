"""CPU, gloo, world_size 2: the N>1 path's host logic — shard the ray range, render each shard, gather.

The per-shard 'renderer' is the oracle (the CUDA engine cannot run here); what is under test is
nerf_b200.sharding: that the shards tile the ray range, that the gathered image equals the
unsharded one bit for bit, and that uniforms keyed on the global ray index make results
shard-invariant.
"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import nerf_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_rays, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nerf_b200 import sharding
    torch.set_num_threads(2)
    sp, sn = O.make_params("proposal", 1, "he"), O.make_params("nerf", 2, "he")
    rays = torch.cat((torch.tensor([0.0, 0.0, 4.0]).expand(n_rays, 3), O.det_uniform((n_rays, 3), 5, -0.3, 0.3) + torch.tensor([0.0, 0.0, -1.0])), -1)
    base_z = torch.linspace(2.0, 6.0, 64)

    def render_rows(start, count):
        ids = list(range(start, start + count))
        jit = O.philox_uniform(99, ids, 64, 0)        # keyed on the GLOBAL ray id
        u = O.philox_uniform(99, ids, 33, 1)
        return O.render_rays(sp, sn, rays[start:start + count], base_z, jit, u, 2.0, 6.0, n_fine=32, white_bkg=True)["rgb"]

    img = sharding.render_image_sharded(render_rows, n_rays)
    # the pre-allocated staging path bench.py uses (nothing allocated per step): same bits, twice in a row
    start, count = sharding.shard_range(n_rays, rank, world)
    buf = sharding.GatherBuffers(n_rays, 3, world, "cpu")
    for _ in range(2):
        again = sharding.gather_rows(img[start:start + count].contiguous(), n_rays, buffers=buf)
        assert torch.equal(again, img)
    if rank == 0:
        q.put(img.clone())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_rank():
    n_rays = 37  # odd on purpose: ragged last shard
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_rays, q)) for r in range(2)]
    for p in procs:
        p.start()
    img2 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    # single "rank": same function, no process group
    sp, sn = O.make_params("proposal", 1, "he"), O.make_params("nerf", 2, "he")
    rays = torch.cat((torch.tensor([0.0, 0.0, 4.0]).expand(n_rays, 3), O.det_uniform((n_rays, 3), 5, -0.3, 0.3) + torch.tensor([0.0, 0.0, -1.0])), -1)
    ids = list(range(n_rays))
    ref = O.render_rays(sp, sn, rays, torch.linspace(2.0, 6.0, 64), O.philox_uniform(99, ids, 64, 0), O.philox_uniform(99, ids, 33, 1),
                        2.0, 6.0, n_fine=32, white_bkg=True)["rgb"]
    assert img2.shape == (n_rays, 3)
    assert float((img2 - ref).abs().max()) < 1e-5


def _grad_worker(rank, world, port, q):
    """ddp_train.py:98 wraps mip_net in DistributedDataParallel, whose one collective is the gradient all-reduce; the engine's
    flat equivalent (train_engine.allreduce_gradients) on two gloo ranks: every rank ends with the average."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import nerf_b200
    from nerf_b200.train_engine import allreduce_gradients
    torch.manual_seed(0)
    net = nerf_b200.MipNeRF(10, 4, 256)
    for i, p in enumerate(net.parameters()):
        p.grad = torch.full_like(p, float(rank + 1)) * (i + 1)
    n = allreduce_gradients([net])
    ok = n == 530052 and all(torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1))) for i, p in enumerate(net.parameters()))
    # and DistributedDataParallel itself accepts the module (the reference's call surface)
    ddp = torch.nn.parallel.DistributedDataParallel(net)
    ok = ok and ddp.module is net
    q.put((rank, bool(ok), n))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in got) and all(n == 530052 for _, _, n in got), got
