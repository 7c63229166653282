"""Mirror of the reference's nerf/mip_model.py: the 8x256 NeRF MLP (called MipNeRF there).

Same constructor, submodule names and state_dict keys as the reference (nerf/mip_model.py:14-38);
forward() runs the whole network — encodings, trunk, skip concat, heads — as one CUDA kernel
(tcgen05 or CUDA-core, per `precision`) through nb2_mlp_forward.
"""
import torch
from torch import nn

from . import _lib, ops
from .nerf_base import NeRF
from .nerf_helper import makeMLP


class MipNeRF(NeRF):
    _nb2_kind = _lib.NET_NERF

    def __init__(self, position_flevel, direction_flevel, hidden_unit=256, cat_origin=True) -> None:
        super().__init__(position_flevel, cat_origin)
        self.direction_flevel = direction_flevel
        self.hidden_unit = hidden_unit
        extra_width = 3 if cat_origin else 0
        module_list = makeMLP(6 * position_flevel + extra_width, hidden_unit)
        for _ in range(3):
            module_list.extend(makeMLP(hidden_unit, hidden_unit))
        self.lin_block1 = nn.Sequential(*module_list)
        self.lin_block2 = nn.Sequential(
            *makeMLP(hidden_unit + 6 * position_flevel + extra_width, hidden_unit),
            *makeMLP(hidden_unit, hidden_unit), *makeMLP(hidden_unit, 256)
        )
        self.bottle_neck = nn.Sequential(*makeMLP(256, 256, None))
        self.opacity_head = nn.Sequential(*makeMLP(256, 1, None))
        self.rgb_layer = nn.Sequential(
            *makeMLP(256 + 6 * direction_flevel + extra_width, 128),
            *makeMLP(128, 3, nn.Sigmoid())
        )
        self.apply(self.init_weight)
        self.precision = None        # fused inference kernels: None -> ops.get_default_precision()
        self.train_precision = None  # differentiable layer-wise engine: None -> 'bf16x3' (or 'bf16')

    def _nb2_linears(self):
        return [self.lin_block1[0], self.lin_block1[2], self.lin_block1[4], self.lin_block1[6],
                self.lin_block2[0], self.lin_block2[2], self.lin_block2[4],
                self.bottle_neck[0], self.opacity_head[0], self.rgb_layer[0], self.rgb_layer[2]]

    def _nb2_levels(self):
        if not self.cat_origin:
            raise _lib.NB2Error("MipNeRF(cat_origin=False) is not supported by the CUDA kernels (the reference never uses it)")
        return self.position_flevel, self.direction_flevel, self.hidden_unit

    def forward(self, pts: torch.Tensor) -> torch.Tensor:
        """pts (ray_num, point_num, 6) = [xyz, dir] -> (ray_num, point_num, 4) = [rgb, sigma]."""
        if self._nb2_wants_grad(pts):
            from .train_engine import NerfEngine, differentiable_forward
            # (a gradient w.r.t. pts covers the positions, columns 0..2; the direction columns receive zeros)
            out = differentiable_forward(self, NerfEngine, _lib.f32(pts if pts.requires_grad else pts.detach()).reshape(-1, pts.shape[-1]), self.train_precision)
            return out.view(pts.shape[0], pts.shape[1], 4)
        net_id = self._nb2_sync()
        out = ops.mlp_forward(net_id, pts.reshape(-1, pts.shape[-1]), self.precision)
        return out.view(pts.shape[0], pts.shape[1], 4)
