"""GPU: every launch variant of the tensor-core MLP kernel computes the same thing.

The default is the CTA-pair kernel (cta_group::2, M256xN256 MMAs, lockstep slots).  The single-CTA kernel
(cta_group::1, N=128 MMAs), its cluster-multicast weight streaming (2 and 4 CTAs) and the ping-pong slot
schedule are selectable through NB2_TC_* environment variables (read at every launch).  All of them perform the
same arithmetic in the same order per tile, so the rendered image must be bit-identical.
"""
import os

import pytest
import torch

import nerf_b200
from nerf_b200 import ops
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
VARIANTS = [
    {"NB2_TC_PAIR": "1", "NB2_TC_LOCKSTEP": "1"},                       # default
    {"NB2_TC_PAIR": "1", "NB2_TC_LOCKSTEP": "0"},
    {"NB2_TC_PAIR": "0", "NB2_TC_CLUSTER": "1", "NB2_TC_LOCKSTEP": "0"},
    {"NB2_TC_PAIR": "0", "NB2_TC_CLUSTER": "1", "NB2_TC_LOCKSTEP": "1"},
    {"NB2_TC_PAIR": "0", "NB2_TC_CLUSTER": "2", "NB2_TC_LOCKSTEP": "1"},
    {"NB2_TC_PAIR": "0", "NB2_TC_CLUSTER": "4", "NB2_TC_LOCKSTEP": "0"},
]


@pytest.mark.parametrize("precision", ["bf16", "fp16x3"])
def test_all_kernel_variants_are_bit_identical(precision):
    prop = nerf_b200.ProposalNetwork(10, 256)
    net = nerf_b200.MipNeRF(10, 4, 256)
    prop.load_state_dict(O.make_params("proposal", 1, "smooth"))
    net.load_state_dict(O.make_params("nerf", 2, "smooth"))
    prop, net = prop.to(DEV), net.to(DEV)
    H = W = 96          # 9216 rays: several tiles per CTA pair, ragged against 148 SMs
    pose = nerf_b200.pose_spherical(75.0, -30.0, 4.0)[:3, :].to(DEV)
    focal = nerf_b200.fov2Focal(0.6911112070083618, (H, W))[0]
    saved = {k: os.environ.get(k) for k in ("NB2_TC_PAIR", "NB2_TC_CLUSTER", "NB2_TC_LOCKSTEP")}
    images = []
    try:
        for v in VARIANTS:
            for k in saved:
                os.environ.pop(k, None)
            os.environ.update(v)
            img = nerf_b200.render_image(net, prop, pose, (H, W), focal, 2.0, 6.0, 128, white_bkg=True, precision=precision, seed=11)["rgb"]
            torch.cuda.synchronize()
            images.append(img.clone())
    finally:
        for k, val in saved.items():
            os.environ.pop(k, None)
            if val is not None:
                os.environ[k] = val
    assert not torch.isnan(images[0]).any() and float(images[0].std()) > 0.01
    for v, img in zip(VARIANTS[1:], images[1:]):
        assert torch.equal(img, images[0]), f"variant {v} differs: max {float((img - images[0]).abs().max())}"
