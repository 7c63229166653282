"""Mirror of the reference's nerf/addtional.py (sic): ProposalNetwork and the PSNR helper.

ProposalNetwork.forward -> nb2_mlp_forward(NB2_NET_PROPOSAL)   (reference nerf/addtional.py:88-96)
ProposalNetwork.get_weights -> nb2_weights_from_sigma           (reference nerf/addtional.py:99-107)
"""
import torch
from torch import nn

from . import _lib, ops
from .nerf_base import PackedModule
from .nerf_helper import makeMLP


def getBounds(weights: torch.Tensor, inds: torch.Tensor):
    """Proposal-weight mass between consecutive fine samples (reference nerf/addtional.py:14-18)."""
    if torch.is_grad_enabled() and weights.requires_grad:
        from .train_engine import GetBounds
        return GetBounds.apply(weights, inds)
    return ops.get_bounds(weights, inds)


class ProposalLoss(nn.Module):
    """reference nerf/addtional.py:20-24 (host-side reduction over CUDA tensors)."""

    def forward(self, prop_bounds: torch.Tensor, nerf_weights: torch.Tensor) -> torch.Tensor:
        bound_diff = (torch.relu(nerf_weights - prop_bounds)) ** 2
        return torch.sum(bound_diff / (nerf_weights + 1e-8))


class SoftL1Loss(nn.Module):
    """Despite the name this is MSE, as in the reference (nerf/addtional.py:38-43)."""

    def __init__(self, epsilon=0.001) -> None:
        super().__init__()
        self.eps = epsilon

    def forward(self, pred: torch.Tensor, target: torch.Tensor):
        return torch.mean((pred - target) ** 2)


class LossPSNR(nn.Module):
    """-10 ln(x) / ln 10 with the reference's constant (nerf/addtional.py:45-51)."""
    __LOG_10__ = 2.3025851249694824

    def forward(self, x):
        return -10. * torch.log(x) / LossPSNR.__LOG_10__


class ProposalNetwork(PackedModule):
    _nb2_kind = _lib.NET_PROPOSAL

    @staticmethod
    def init_weight(m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)

    def __init__(self, position_flevel, hidden_unit=128, cat_origin=True) -> None:
        super().__init__()
        self.position_dims = position_flevel * 6
        self.position_flevel = position_flevel
        self.cat_origin = cat_origin
        self.hidden_unit = hidden_unit
        extra_dims = 3 if cat_origin else 0
        self.layers = nn.Sequential(
            *makeMLP(self.position_dims + extra_dims, hidden_unit),
            *makeMLP(hidden_unit, hidden_unit), *makeMLP(hidden_unit, hidden_unit), *makeMLP(hidden_unit, hidden_unit),
            *makeMLP(hidden_unit, 1, None)
        )
        self.apply(self.init_weight)
        self.precision = None
        self.train_precision = None

    def loadFromFile(self, load_path: str, use_amp=False, other_stuff=None):
        save = torch.load(load_path, map_location="cpu")
        save_model = save['model']
        state_dict = {k: save_model[k] for k in self.state_dict().keys()}
        model_dict = self.state_dict()
        model_dict.update(state_dict)
        self.load_state_dict(model_dict)
        if use_amp:
            raise _lib.NB2Error("apex amp state is not supported by nerf_b200 (precision is chosen per call)")
        print("NeRF Model loaded from '%s'" % (load_path))
        if other_stuff is not None:
            return [save[k] for k in other_stuff]

    def _nb2_linears(self):
        return [self.layers[0], self.layers[2], self.layers[4], self.layers[6], self.layers[8]]

    def _nb2_levels(self):
        if not self.cat_origin:
            raise _lib.NB2Error("ProposalNetwork(cat_origin=False) is not supported by the CUDA kernels")
        return self.position_flevel, 0, self.hidden_unit

    def forward(self, pts: torch.Tensor, encoded_pt: torch.Tensor = None) -> torch.Tensor:
        """pts (ray_num, point_num, 3) -> raw density (ray_num, point_num)."""
        if self._nb2_wants_grad(pts, encoded_pt):
            if encoded_pt is not None:
                raise _lib.NB2Error("ProposalNetwork.forward(encoded_pt=...) is inference-only (the reference never trains through it)")
            from .train_engine import ProposalEngine, differentiable_forward
            # positions keep their graph when they ask for a gradient (train.py:165-168, --prop_normal: get_grad(density, samples))
            out = differentiable_forward(self, ProposalEngine, _lib.f32(pts if pts.requires_grad else pts.detach()).reshape(-1, 3), self.train_precision)
            return out.view(pts.shape[0], pts.shape[1])
        net_id = self._nb2_sync()
        if encoded_pt is not None:
            # the reference views encoded_pt as (R, P, position_dims) and concatenates it behind the raw points
            enc = encoded_pt.reshape(-1, self.position_dims)
            out = ops.mlp_forward_encoded(net_id, pts.reshape(-1, 3), enc, self.precision)
            return out.view(pts.shape[0], pts.shape[1])
        out = ops.mlp_forward(net_id, pts.reshape(-1, 3), self.precision)
        return out.view(pts.shape[0], pts.shape[1])

    @staticmethod
    def get_weights(density: torch.Tensor, zvals: torch.Tensor, ray_dirs: torch.Tensor = None) -> torch.Tensor:
        if torch.is_grad_enabled() and density.requires_grad:
            from .train_engine import WeightsFromSigma
            return WeightsFromSigma.apply(density, zvals.detach(), ray_dirs.detach() if ray_dirs is not None else None, "relu")
        return ops.weights_from_sigma(density, zvals, ray_dirs, "relu")
