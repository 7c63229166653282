"""Mirror of the reference's nerf/param_com.py (model_average.py:238-251): parameter exchange between the ranks of a
model-averaging run.  Same names, arguments and results; the reference issues one collective / point-to-point call PER
PARAMETER TENSOR (22 for MipNeRF, 10 for the proposal network), here every function moves ONE flat buffer (2.1 MB for
MipNeRF) -- over NCCL / NVLink that is one launch instead of 22.  Element-wise the arithmetic is the same: send / recv /
recv_avg give bit-identical weights to the per-tensor version; reduce / all_reduce agree to the summation order of the
collective (a ring sums an element in the order of the chunk it falls in), i.e. a few ulps with three or more ranks
(tests/test_distributed_cpu.py runs both versions over gloo on three ranks).
Both ends of a send / recv pair must use this module (the message is the flat buffer)."""
from typing import List

import torch
from torch import distributed as dist


def _params(model):
    return [p.data for p in model.parameters()]


def _flatten(tensors):
    return torch.cat([t.reshape(-1) for t in tensors]) if tensors else torch.empty(0)


def _scatter_back(flat, tensors):
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n


def param_send(model, dist_ranks: List[int], group=None):
    """Send every parameter to the listed ranks (param_com.py:13-17)."""
    flat = _flatten(_params(model))
    for rank in dist_ranks:
        dist.send(tensor=flat, dst=rank, group=group)


def param_recv(model, source_rank: int, group=None):
    """Receive every parameter from `source_rank` (param_com.py:19-22)."""
    ps = _params(model)
    flat = _flatten(ps)
    dist.recv(tensor=flat, src=source_rank, group=group)
    _scatter_back(flat, ps)


def param_recv_avg(model, tmp, weights: list, source_ranks: List[int], self_rank: int = 0, group=None):
    """model = weights[self] * model + sum over sources of weights[src] * received (param_com.py:24-34); `tmp` ends up
    holding the parameters of the last source, as in the reference."""
    ps = _params(model)
    flat = _flatten(ps)
    flat *= weights[self_rank]
    buf = torch.empty_like(flat)
    for src_rank in source_ranks:
        dist.recv(tensor=buf, src=src_rank, group=group)
        flat += weights[src_rank] * buf
    _scatter_back(flat, ps)
    if source_ranks:
        _scatter_back(buf, _params(tmp))


def param_reduce(model, weights: list, self_rank: int, dst_rank: int = 0, group=None):
    """Scale by this rank's weight, then reduce (sum) onto `dst_rank` (param_com.py:36-42)."""
    ps = _params(model)
    flat = _flatten(ps)
    flat *= weights[self_rank]
    dist.reduce(tensor=flat, dst=dst_rank, group=group)
    _scatter_back(flat, ps)


def param_broadcast(model, src_rank: int = 0, group=None):
    """Broadcast every parameter from `src_rank` (param_com.py:44-47)."""
    ps = _params(model)
    flat = _flatten(ps)
    dist.broadcast(tensor=flat, src=src_rank, group=group)
    _scatter_back(flat, ps)


def param_all_reduce(model, group=None):
    """One-step model average: the parameters are already weighted by the caller (param_com.py:49-54)."""
    ps = _params(model)
    flat = _flatten(ps)
    dist.all_reduce(tensor=flat, group=group)
    _scatter_back(flat, ps)
