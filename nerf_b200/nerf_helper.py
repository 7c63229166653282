"""Mirror of the reference's nerf/nerf_helper.py for the hot path.

positional_encoding -> CUDA kernel `posenc_kernel` via nb2_posenc
(reference: /root/reference/nerf/nerf_helper.py:38-48).
"""
import torch

from . import ops


def makeMLP(in_chan, out_chan, act=torch.nn.ReLU(), batch_norm=False):
    """Same module list as the reference (nerf_helper.py:17-23) so state_dict keys line up."""
    modules = [torch.nn.Linear(in_chan, out_chan)]
    if batch_norm:
        modules.append(torch.nn.BatchNorm1d(out_chan))
    if act is not None:
        modules.append(act)
    return modules


def saveModel(model, path, other_stuff=None, opt=None, amp=None):
    """Checkpoint format of the reference (nerf_helper.py:7-15)."""
    checkpoint = {"model": model.state_dict()}
    if amp is not None:
        checkpoint["amp"] = amp.state_dict()
    if opt is not None:
        checkpoint["optimizer"] = opt.state_dict()
    if other_stuff is not None:
        checkpoint.update(other_stuff)
    torch.save(checkpoint, path)


def positional_encoding(x, freq_level):
    """[sin(2^f x), cos(2^f x)] for f < freq_level, concatenated on the last axis.

    x: (..., 3) CUDA tensor.  Returns (ray_num, point_num, 6L) for 3-D input, (n, 6L) otherwise
    (the reference's view logic, nerf_helper.py:45-47).
    """
    enc = ops.posenc(x, freq_level)
    if x.dim() > 2:
        return enc.view(x.shape[0], x.shape[1], -1)
    return enc.view(*x.shape[:-1], -1)


def linear_to_srgb(linear, eps=None):
    """Reference nerf/nerf_helper.py:50-56 (eps is torch.finfo(float32).eps there; other values are not supported)."""
    if eps is not None:
        raise ValueError("linear_to_srgb: only the reference's default eps is implemented")
    return ops.linear_to_srgb(linear)
