"""nerf_b200 — B200 (sm_100a) engine for the ray-marching hot path of Enigmatisms/NeRF.

The package mirrors the reference's `nerf` package for that path (same module, class and
function names): nerf_helper, nerf_base, mip_model, addtional, mip_methods, utils, procedures.
All numerical work happens in libnerfb200.so (hand-written CUDA behind the C ABI declared in
include/nerf_b200.h); there is no CPU or PyTorch fallback.
"""
from . import _lib, ops  # noqa: F401
from ._lib import NB2Error  # noqa: F401
from .ops import get_default_precision, set_default_precision  # noqa: F401
from .nerf_helper import makeMLP, positional_encoding, saveModel, linear_to_srgb  # noqa: F401
from .ref_func import generate_ide_fn  # noqa: F401
from .nerf_base import NeRF, DecayLrScheduler  # noqa: F401
from .mip_model import MipNeRF  # noqa: F401
from .ref_model import BackFaceLoss, RefNeRF, WeightedNormalLoss  # noqa: F401
from . import param_com  # noqa: F401
from .addtional import ProposalNetwork, LossPSNR, SoftL1Loss, ProposalLoss, getBounds  # noqa: F401
from .mip_methods import maxBlurFilter, ipe_feature  # noqa: F401
from .utils import inverseSample, sample_pdf, fov2Focal, pose_spherical, validSampler  # noqa: F401
from .procedures import render_image, get_patch_size, get_parser, render_only  # noqa: F401

__version__ = "0.1.0"
