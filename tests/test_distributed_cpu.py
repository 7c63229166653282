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


def _per_tensor_param_com():
    """The reference's per-parameter exchange (nerf/param_com.py:13-54), restated as the oracle of the flat versions."""
    def send(model, ranks, group=None):
        for p in model.parameters():
            for r in ranks:
                dist.send(tensor=p.data, dst=r, group=group)

    def recv(model, src, group=None):
        for p in model.parameters():
            dist.recv(tensor=p.data, src=src, group=group)

    def recv_avg(model, tmp, weights, srcs, self_rank=0, group=None):
        for p_tmp, p_model in zip(tmp.parameters(), model.parameters()):
            p_model.data *= weights[self_rank]
            for s in srcs:
                dist.recv(tensor=p_tmp.data, src=s, group=group)
                p_model.data += weights[s] * p_tmp.data

    def reduce(model, weights, self_rank, dst=0, group=None):
        for p in model.parameters():
            p.data *= weights[self_rank]
            dist.reduce(tensor=p.data, dst=dst, group=group)

    def broadcast(model, src=0, group=None):
        for p in model.parameters():
            dist.broadcast(tensor=p.data, src=src, group=group)

    def all_reduce(model, group=None):
        for p in model.parameters():
            dist.all_reduce(tensor=p.data, group=group)
    return send, recv, recv_avg, reduce, broadcast, all_reduce


def _param_com_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import copy
    import nerf_b200
    from nerf_b200 import param_com as pc
    r_send, r_recv, r_recv_avg, r_reduce, r_broadcast, r_all_reduce = _per_tensor_param_com()
    torch.manual_seed(100 + rank)
    base = nerf_b200.ProposalNetwork(10, 256)
    for p in base.parameters():
        p.data.uniform_(-1.0, 1.0)
    weights = [0.5, 0.3, 0.2]
    others = [r for r in range(world) if r != 0]
    results = []

    def scenario(impl):
        send, recv, recv_avg, reduce, broadcast, all_reduce = impl
        out = []
        # model_average.py:236-244: rank 0 averages what the others send, then sends the average back
        m, tmp = copy.deepcopy(base), copy.deepcopy(base)
        if rank == 0:
            recv_avg(m, tmp, weights, others, 0)
            send(m, others)
        else:
            send(m, [0])
            recv(m, 0)
        out.append([p.data.clone() for p in m.parameters()] + ([p.data.clone() for p in tmp.parameters()] if rank == 0 else []))
        # :246-247: weighted reduce onto rank 0, broadcast back
        m = copy.deepcopy(base)
        reduce(m, weights, rank, 0)
        broadcast(m, 0)
        out.append([p.data.clone() for p in m.parameters()])
        # :249-251: pre-weighted all-reduce
        m = copy.deepcopy(base)
        for p in m.parameters():
            p.data *= weights[rank]
        all_reduce(m)
        out.append([p.data.clone() for p in m.parameters()])
        return out
    ours = scenario((pc.param_send, pc.param_recv, pc.param_recv_avg, pc.param_reduce, pc.param_broadcast, pc.param_all_reduce))
    dist.barrier()
    ref = scenario((r_send, r_recv, r_recv_avg, r_reduce, r_broadcast, r_all_reduce))
    # point-to-point exchange + weighted average: the same element-wise arithmetic in the same order -> bit-identical;
    # reduce / all-reduce: a ring sums an element in the order of the chunk it falls in, which moves with the buffer layout
    # (3+ ranks) -> equal to a few ulps
    same = all(len(sa) == len(sb) for sa, sb in zip(ours, ref)) and all(torch.equal(a, b) for a, b in zip(ours[0], ref[0]))
    same = same and all(float((a - b).abs().max()) <= 4e-7 for s in (1, 2) for a, b in zip(ours[s], ref[s]))
    # every rank ends each scenario with the same averaged model
    digest = [float(sum(t.double().sum() for t in s[:10])) for s in ours]
    q.put((rank, bool(same), digest))
    dist.barrier()
    dist.destroy_process_group()


def test_param_com_flat_exchange_equals_per_tensor_exchange():
    """nerf_b200.param_com (one flat buffer per call) against the reference's per-parameter loops on three gloo ranks."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_param_com_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=300) for _ in range(3))
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert all(same for _, same, _ in got), got
    for s in range(3):
        assert got[0][2][s] == got[1][2][s] == got[2][2][s], got       # every rank holds the same averaged model
