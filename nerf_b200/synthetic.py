"""Synthetic workloads: deterministic, platform-independent parameters and uniforms for benchmarks and tests.

There is no dataset or checkpoint in the build / bench environment, so bench.py, smoke() and the tests run on
random-init weights of the reference architectures.  Everything here is generated from integer hashes with exact
fp32 conversions, so the GPU box and the build container see bit-identical inputs without shipping them.  This module
is data generation only (numpy on the host); it is neither the engine nor the oracle, and both may import it.
"""
import math

import numpy as np
import torch

_MASK = (1 << 64) - 1


def _mix64(x):
    """splitmix64 finaliser on a numpy uint64 array."""
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(_MASK)
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(_MASK)
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & np.uint64(_MASK)
    return x ^ (x >> np.uint64(31))


def det_uniform(shape, seed, lo=-1.0, hi=1.0):
    """Deterministic fp32 uniforms in [lo, hi): 24-bit hash -> exact float32 arithmetic."""
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x100000001B3)
        bits = _mix64(idx) >> np.uint64(40)
    u = bits.astype(np.float32) * np.float32(1.0 / 16777216.0)
    out = (np.float32(lo) + u * np.float32(hi - lo)).astype(np.float32)
    return torch.from_numpy(out.reshape(shape))


PROPOSAL_KEYS = ["layers.0", "layers.2", "layers.4", "layers.6", "layers.8"]
NERF_KEYS = ["lin_block1.0", "lin_block1.2", "lin_block1.4", "lin_block1.6", "lin_block2.0", "lin_block2.2",
             "lin_block2.4", "bottle_neck.0", "opacity_head.0", "rgb_layer.0", "rgb_layer.2"]


def layer_shapes(kind, pos_levels=10, dir_levels=4, hidden=256):
    """(out, in) of every nn.Linear, in state_dict order (nerf/addtional.py:67-71, nerf/mip_model.py:19-37)."""
    enc, denc = 3 + 6 * pos_levels, 3 + 6 * dir_levels
    if kind == "proposal":
        return [(hidden, enc), (hidden, hidden), (hidden, hidden), (hidden, hidden), (1, hidden)]
    return [(hidden, enc), (hidden, hidden), (hidden, hidden), (hidden, hidden), (hidden, hidden + enc),
            (hidden, hidden), (256, hidden), (256, 256), (1, 256), (128, 256 + denc), (3, 128)]


def np_forward(kind, sd, pts, pos_levels=10, dir_levels=4):
    """fp64 numpy forward of either MLP (same citations as proposal_forward / nerf_forward below).
    Used to calibrate the synthetic density head and as the "infinitely precise" yardstick when the
    tests compare the error of the engine with the error of the reference's own fp32 arithmetic."""
    g = {k: v.numpy().astype(np.float64) for k, v in sd.items()}
    pts = np.asarray(pts, dtype=np.float64)

    def enc(x, levels):
        parts = [x]
        for f in range(levels):
            parts += [np.sin((2.0 ** f) * x), np.cos((2.0 ** f) * x)]
        return np.concatenate(parts, axis=-1)

    def lin(h, key, relu=True):
        y = h @ g[key + ".weight"].T + g[key + ".bias"]
        return np.maximum(y, 0.0) if relu else y

    x = pts[..., :3]
    ex = enc(x, pos_levels)
    if kind == "proposal":
        h = ex
        for key in PROPOSAL_KEYS[:-1]:
            h = lin(h, key)
        return lin(h, "layers.8", relu=False)[..., 0]
    d = pts[..., 3:6]
    er = enc(d / np.linalg.norm(d, axis=-1, keepdims=True), dir_levels)
    h = ex
    for key in NERF_KEYS[:4]:
        h = lin(h, key)
    h = np.concatenate((ex, h), axis=-1)
    for key in NERF_KEYS[4:7]:
        h = lin(h, key)
    sigma = lin(h, "opacity_head.0", relu=False)
    b = lin(h, "bottle_neck.0", relu=False)
    t = lin(np.concatenate((b, er), axis=-1), "rgb_layer.0")
    rgb = 1.0 / (1.0 + np.exp(-lin(t, "rgb_layer.2", relu=False)))
    return np.concatenate((rgb, sigma), axis=-1)


def make_params(kind, seed, style="he", pos_levels=10, dir_levels=4, hidden=256, sigma_std=25.0, sigma_mean=-15.0):
    """state_dict-shaped deterministic parameters.

    style 'he'      : uniform with He variance and small biases; the density head is then rescaled and
                      re-biased (from an fp64 forward over fixed probe points) so the raw density has
                      mean `sigma_mean` and std `sigma_std` over the scene volume -> a non-degenerate
                      random field with empty and opaque regions.  (A deep ReLU net at init is almost
                      constant in space, so without this every sample would have the same sign.)
    style 'smooth'  : as 'he', but the first-layer / skip-layer columns that read encoding level l are damped
                      by 2^-l.  'he' is a *chaotic* field: its outputs move by ~1e-4 when a sample depth moves by
                      one fp32 ulp (the 2^9 encoding frequency times an O(1) gain), so no two implementations
                      of the reference (not even its own CPU and GPU paths) agree to 1e-4 end to end on it.
                      'smooth' is band-limited the way trained radiance fields are, and is the field used for
                      end-to-end parity.
    style 'refinit' : the reference's init scale (std 0.02, zero bias; nerf/nerf_base.py:14-22) ->
                      the tiny-activation regime every freshly constructed reference model is in.
    """
    keys = PROPOSAL_KEYS if kind == "proposal" else NERF_KEYS
    head = "layers.8" if kind == "proposal" else "opacity_head.0"
    sd = {}
    for i, (key, (o, k)) in enumerate(zip(keys, layer_shapes(kind, pos_levels, dir_levels, hidden))):
        if style in ("he", "smooth"):
            bound = math.sqrt(6.0 / k)
            w = det_uniform((o, k), seed * 1000 + 2 * i, -bound, bound)
            b = det_uniform((o,), seed * 1000 + 2 * i + 1, -0.1, 0.1)
        else:
            bound = 0.02 * math.sqrt(3.0)
            w = det_uniform((o, k), seed * 1000 + 2 * i, -bound, bound)
            b = torch.zeros(o)
        if style == "smooth" and key in ("layers.0", "lin_block1.0", "lin_block2.0"):
            # band-limit the field like a trained network: the columns fed by encoding level l are damped
            # by 2^-l, so d(output)/d(position) stays O(1) instead of O(2^9)
            damp = np.ones(k, dtype=np.float32)
            for l in range(pos_levels):
                damp[3 + 6 * l: 9 + 6 * l] = np.float32(2.0 ** -l)
            damp[:3 + 6 * pos_levels] *= np.float32(2.5)
            w = torch.from_numpy((w.numpy() * damp[None, :]).astype(np.float32))
        sd[key + ".weight"], sd[key + ".bias"] = w, b
    if style in ("he", "smooth"):
        probe = torch.cat((det_uniform((1024, 3), seed * 1000 + 777, -2.0, 2.0),
                           det_uniform((1024, 3), seed * 1000 + 778, -1.0, 1.0)), dim=-1).numpy()
        out = np_forward(kind, sd, probe, pos_levels, dir_levels)
        sig = out if kind == "proposal" else out[..., 3]
        gain = sigma_std / float(sig.std())
        # round the calibration constants to 6 significant digits: robust to last-bit fp64 differences
        gain = float(f"{gain:.6g}")
        shift = float(f"{sigma_mean - gain * float(sig.mean() - sd[head + '.bias'].item()):.6g}")
        sd[head + ".weight"] = (sd[head + ".weight"].numpy() * np.float32(gain)).astype(np.float32)
        sd[head + ".weight"] = torch.from_numpy(sd[head + ".weight"])
        sd[head + ".bias"] = torch.tensor([shift], dtype=torch.float32)
    return sd


def params_to(sd, device=None, dtype=None):
    return {k: v.to(device=device, dtype=dtype) for k, v in sd.items()}




def det_state_dict(module, seed, gain=1.0, bias=0.05):
    """Deterministic He-uniform parameters for ANY module made of nn.Linear layers (used for Ref-NeRF, whose layer
    table differs from the two networks above): weight ~ U(+-gain * sqrt(6 / fan_in)), bias ~ U(+-bias)."""
    sd = {}
    for i, (k, v) in enumerate(module.state_dict().items()):
        if k.endswith(".weight"):
            bound = gain * math.sqrt(6.0 / v.shape[1])
            sd[k] = det_uniform(tuple(v.shape), seed * 1000 + i, -bound, bound)
        else:
            sd[k] = det_uniform(tuple(v.shape), seed * 1000 + i, -bias, bias)
    return sd
