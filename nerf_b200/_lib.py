"""ctypes binding of libnerfb200.so (include/nerf_b200.h).

The shared library is the product: there is no Python/CPU fallback.  Importing this module
never touches the GPU; the first call that needs the library loads it and fails loudly if it
(or a CUDA device) is missing.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# NB2_LIB selects another build of the same library (e.g. libnerfb200_prof.so: role cycle counters compiled in)
LIB_PATH = os.path.join(_HERE, os.environ.get("NB2_LIB", "libnerfb200.so"))

NET_PROPOSAL, NET_NERF = 0, 1
PREC_FP32, PREC_FP16X3, PREC_BF16, PREC_FP16, PREC_BF16X3 = 0, 1, 2, 3, 4
PRECISIONS = {"fp32": PREC_FP32, "fp16x3": PREC_FP16X3, "bf16": PREC_BF16, "fp16": PREC_FP16, "bf16x3": PREC_BF16X3,
              "fp16m": 5, "bf16m": 6}     # mixed (render_rays only): proposal network split precision, NeRF network single pass
WHITE_BKG, DENSITY_SOFTPLUS, PROPOSAL_IPE = 1, 2, 4

c_f32p = ctypes.c_void_p
c_i64 = ctypes.c_int64
c_u64 = ctypes.c_uint64
c_int = ctypes.c_int
c_float = ctypes.c_float
c_vp = ctypes.c_void_p


class RenderParams(ctypes.Structure):
    """Mirror of nb2_render_params."""
    _fields_ = [
        ("n_coarse", c_int), ("n_fine", c_int),
        ("near_t", c_float), ("far_t", c_float),
        ("resolution", c_float), ("blur_alpha", c_float),
        ("flags", c_int), ("precision", c_int),
        ("seed", c_u64), ("ray_offset", c_i64),
        ("prop_net_id", c_int), ("nerf_net_id", c_int),
        ("n_peers", c_int), ("ipe_radius", c_float),
        ("peer_rgb", c_vp * 8),
    ]


class GemmOperand(ctypes.Structure):
    """Mirror of nb2_gemm_operand."""
    _fields_ = [("ptr", c_vp), ("ld", c_i64), ("mn_major", c_int), ("reserved", c_int)]


class GemmSeg(ctypes.Structure):
    _fields_ = [("a", GemmOperand), ("b", GemmOperand), ("K", c_i64)]


GEMM_MAX_SEG = 8


class GemmDesc(ctypes.Structure):
    """Mirror of nb2_gemm_desc."""
    _fields_ = [
        ("M", c_i64), ("N", c_int), ("n_seg", c_int),
        ("seg", GemmSeg * GEMM_MAX_SEG),
        ("bias", c_vp), ("act", c_int), ("ld_mask", c_int), ("mask", c_vp),
        ("out_f32", c_vp), ("out_hi", c_vp), ("out_lo", c_vp),
        ("ld_f32", c_i64), ("ld_16", c_i64),
        ("splits", c_int), ("reserved", c_int), ("split_stride", c_i64),
        ("a_rowsum_out", c_vp), ("a_rowsum_stride", c_i64),
    ]


class ToBf16Desc(ctypes.Structure):
    """Mirror of nb2_to_bf16_desc."""
    _fields_ = [("src", c_vp), ("rows", c_i64), ("ld_src", c_i64), ("cols", c_int), ("ld_dst", c_int), ("col_perm", c_vp), ("hi", c_vp), ("lo", c_vp)]


class ReduceDesc(ctypes.Structure):
    """Mirror of nb2_reduce_desc."""
    _fields_ = [
        ("ws", c_vp), ("splits", c_int), ("rows", c_int), ("cols", c_int), ("ld_ws", c_int), ("ld_out", c_int), ("accumulate", c_int),
        ("split_stride", c_i64), ("col_perm", c_vp), ("out", c_vp),
    ]


# name -> (restype, argtypes).  Every symbol include/nerf_b200.h declares is listed here; the
# CPU test-suite checks the library exports all of them.
SIGNATURES = {
    "nb2_last_error": (ctypes.c_char_p, []),
    "nb2_version": (c_int, []),
    "nb2_create": (c_int, [ctypes.POINTER(c_vp), c_int]),
    "nb2_destroy": (c_int, [c_vp]),
    "nb2_pack_weights": (c_int, [c_vp, c_int, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), c_int, c_int, c_int, c_int, c_vp]),
    "nb2_weights_version": (c_int, [c_vp, c_int]),
    "nb2_net_create": (c_int, [c_vp, c_int, ctypes.POINTER(c_int)]),
    "nb2_net_destroy": (c_int, [c_vp, c_int]),
    "nb2_gemm_bf16": (c_int, [c_vp, ctypes.POINTER(GemmDesc), c_vp]),
    "nb2_to_bf16": (c_int, [c_vp, c_f32p, c_i64, c_int, c_i64, c_vp, c_vp, c_vp, c_int, c_vp]),
    "nb2_to_bf16_batch": (c_int, [c_vp, c_vp, c_int, c_vp]),
    "nb2_reduce_splits": (c_int, [c_vp, c_f32p, c_int, c_i64, c_int, c_int, c_int, c_vp, c_f32p, c_int, c_int, c_vp]),
    "nb2_gemm_bf16_batch": (c_int, [c_vp, c_vp, c_int, c_vp]),
    "nb2_reduce_splits_batch": (c_int, [c_vp, c_vp, c_int, c_vp]),
    "nb2_encode_bf16": (c_int, [c_vp, c_f32p, c_int, c_int, c_i64, c_int, c_int, c_vp, c_vp, c_i64, c_int, c_vp]),
    "nb2_weights_from_sigma_backward": (c_int, [c_vp, c_f32p, c_f32p, c_f32p, c_int, c_i64, c_int, c_int, c_f32p, c_f32p, c_vp]),
    "nb2_composite_backward": (c_int, [c_vp, c_f32p, c_f32p, c_f32p, c_int, c_i64, c_int, c_int, c_f32p, c_f32p, c_f32p, c_vp]),
    "nb2_max_blur_backward": (c_int, [c_vp, c_f32p, c_f32p, c_i64, c_int, c_f32p, c_vp]),
    "nb2_get_bounds_backward": (c_int, [c_vp, c_vp, c_f32p, c_i64, c_int, c_int, c_f32p, c_vp]),
    "nb2_nerf_head_backward": (c_int, [c_vp, c_f32p, c_f32p, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "nb2_ref_color_backward": (c_int, [c_vp, c_f32p, c_f32p, c_int, c_int, c_f32p, c_i64, c_vp, c_vp, c_f32p, c_int, c_vp]),
    "nb2_ref_geometry_backward": (c_int, [c_vp, c_f32p, c_int, c_f32p, c_int, c_i64, c_f32p, c_i64, c_int, c_f32p, c_f32p, c_vp, c_int, c_int,
                                          c_f32p, c_int, c_vp]),
    "nb2_encode_backward": (c_int, [c_vp, c_f32p, c_int, c_int, c_i64, c_int, c_f32p, c_i64, c_f32p, c_vp]),
    "nb2_dot3": (c_int, [c_vp, c_f32p, c_f32p, c_i64, c_f32p, c_vp]),
    "nb2_ref_geometry": (c_int, [c_vp, c_f32p, c_int, c_f32p, c_int, c_i64, c_f32p, c_f32p, c_f32p, c_f32p, c_vp]),
    "nb2_ref_dir_inputs": (c_int, [c_vp, c_f32p, c_int, c_f32p, c_i64, c_vp, c_vp, c_i64, c_int, c_vp]),
    "nb2_ref_color": (c_int, [c_vp, c_f32p, c_f32p, c_int, c_int, c_int, c_f32p, c_f32p, c_i64, c_f32p, c_f32p, c_vp]),
    "nb2_coarse_fine_merge_inds": (c_int, [c_vp, c_f32p, c_f32p, c_f32p, c_vp, c_i64, c_int, c_int, c_f32p, c_f32p, c_vp, c_vp, c_vp]),
    "nb2_composite_aux": (c_int, [c_vp, c_f32p, c_f32p, c_f32p, c_int, c_i64, c_int, c_int, c_float, c_float, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p,
                                  c_f32p, c_vp]),
    "nb2_ipc_alloc": (c_int, [c_vp, c_i64, ctypes.POINTER(c_vp), c_vp]),
    "nb2_ipc_open": (c_int, [c_vp, c_vp, ctypes.POINTER(c_vp)]),
    "nb2_ipc_close": (c_int, [c_vp, c_vp]),
    "nb2_ipc_free": (c_int, [c_vp, c_vp]),
    "nb2_generate_rays": (c_int, [c_vp, c_f32p, c_int, c_int, c_float, c_float, c_i64, c_i64, c_f32p, c_vp]),
    "nb2_sample_coarse": (c_int, [c_vp, c_f32p, c_f32p, c_f32p, c_float, c_u64, c_i64, c_i64, c_int, c_f32p, c_f32p, c_vp]),
    "nb2_posenc": (c_int, [c_vp, c_f32p, c_i64, c_int, c_int, c_f32p, c_vp]),
    "nb2_ipe": (c_int, [c_vp, c_f32p, c_f32p, c_i64, c_int, c_int, c_float, c_f32p, c_f32p, c_f32p, c_vp, c_vp]),
    "nb2_weights_from_sigma": (c_int, [c_vp, c_f32p, c_f32p, c_f32p, c_int, c_i64, c_int, c_int, c_f32p, c_vp]),
    "nb2_max_blur": (c_int, [c_vp, c_f32p, c_i64, c_int, c_float, c_f32p, c_vp]),
    "nb2_sample_pdf": (c_int, [c_vp, c_f32p, c_f32p, c_f32p, c_u64, c_i64, c_i64, c_int, c_int, c_f32p, c_vp, c_vp, c_vp]),
    "nb2_inverse_sample": (c_int, [c_vp, c_f32p, c_f32p, c_f32p, c_u64, c_i64, c_i64, c_int, c_int, c_int, c_f32p, c_vp, c_vp]),
    "nb2_search_cdf": (c_int, [c_vp, c_f32p, c_f32p, c_i64, c_int, c_int, c_vp, c_vp]),
    "nb2_resample": (c_int, [c_vp, c_f32p, c_f32p, c_f32p, c_f32p, c_u64, c_i64, c_i64, c_int, c_int, c_float, c_int, c_f32p, c_vp, c_vp]),
    "nb2_length2pts": (c_int, [c_vp, c_f32p, c_f32p, c_i64, c_int, c_f32p, c_vp]),
    "nb2_coarse_fine_merge": (c_int, [c_vp, c_f32p, c_f32p, c_f32p, c_i64, c_int, c_int, c_f32p, c_f32p, c_vp]),
    "nb2_valid_sampler": (c_int, [c_vp, c_f32p, c_vp, c_f32p, c_vp, c_f32p, c_f32p, c_float, c_float, c_float, c_u64, c_i64, c_i64, c_i64,
                                  c_int, c_f32p, c_f32p, c_f32p, c_f32p, c_vp]),
    "nb2_get_bounds": (c_int, [c_vp, c_f32p, c_vp, c_i64, c_int, c_int, c_f32p, c_vp]),
    "nb2_ide": (c_int, [c_vp, c_f32p, c_f32p, c_i64, c_f32p, c_vp, c_int, c_int, c_f32p, c_vp]),
    "nb2_linear_to_srgb": (c_int, [c_vp, c_f32p, c_i64, c_f32p, c_vp]),
    "nb2_mlp_forward": (c_int, [c_vp, c_int, c_int, c_f32p, c_int, c_i64, c_f32p, c_vp]),
    "nb2_mlp_forward_encoded": (c_int, [c_vp, c_int, c_int, c_f32p, c_int, c_f32p, c_i64, c_f32p, c_vp]),
    "nb2_composite": (c_int, [c_vp, c_f32p, c_f32p, c_f32p, c_int, c_i64, c_int, c_int, c_float, c_float, c_f32p, c_f32p, c_f32p, c_f32p, c_vp]),
    "nb2_render_workspace_bytes": (c_i64, [c_i64, ctypes.POINTER(RenderParams)]),
    "nb2_render_rays": (c_int, [c_vp, ctypes.POINTER(RenderParams), c_f32p, c_f32p, c_f32p, c_f32p, c_i64, c_f32p, c_f32p,
                                c_f32p, c_f32p, c_f32p, c_f32p, c_vp, c_vp, c_i64, c_vp]),
    "nb2_launch_count": (c_i64, [c_vp]),
    "nb2_set_profile_events": (c_int, [c_vp, ctypes.POINTER(c_vp)]),
    "nb2_debug_tc_profile": (c_int, [c_vp, ctypes.POINTER(ctypes.c_longlong), c_int]),
    "nb2_debug_umma_bench": (c_int, [c_vp, c_vp, c_vp, c_f32p, c_vp, c_int, c_int, c_int, c_vp, c_vp]),
    "nb2_debug_microbench": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_int, c_vp, ctypes.POINTER(c_int), c_vp]),
    "nb2_selftest_umma": (c_int, [c_vp, c_vp, c_vp, c_vp, c_f32p, c_vp]),
    "nb2_selftest_umma_ts": (c_int, [c_vp, c_vp, c_vp, c_vp, c_f32p, c_vp]),
}

_lib = None
_lib_lock = threading.Lock()
_handles = {}


class NB2Error(RuntimeError):
    pass


def load():
    """Load libnerfb200.so (no GPU needed for the load itself)."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NB2Error(
                f"{LIB_PATH} is missing: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
                "nerf_b200 has no CPU or PyTorch fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc):
    if rc != 0:
        msg = load().nb2_last_error()
        raise NB2Error(f"libnerfb200 error {rc}: {msg.decode() if msg else '?'}")


def handle(device=None):
    """One nb2_handle per CUDA device per process."""
    if device is not None:
        try:
            return _handles_fast[device]              # a torch.device with an index (what tensors carry): no torch calls
        except (KeyError, TypeError):
            pass
    if not torch.cuda.is_available():
        raise NB2Error("nerf_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    idx = _device_index(device)
    h = _handles.get(idx)
    if h is None:
        lib = load()
        out = c_vp()
        torch.cuda.init()
        check(lib.nb2_create(ctypes.byref(out), idx))
        h = out
        _handles[idx] = h
    if isinstance(device, torch.device) and device.index is not None:
        _handles_fast[device] = h
    return h


_handles_fast = {}
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _device_index(device):
    if device is None:
        return torch.cuda.current_device()
    idx = torch.device(device).index if not isinstance(device, int) else device
    return torch.cuda.current_device() if idx is None else idx


def stream_ptr(device=None):
    """The current CUDA stream of `device` (default: the current device).  Ops pass the device of their tensors, so a
    tensor on a non-current GPU is processed on that GPU's stream by that GPU's handle."""
    if _raw_stream is not None:
        idx = device.index if isinstance(device, torch.device) and device.index is not None else _device_index(device)
        return c_vp(_raw_stream(idx))
    return c_vp(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise NB2Error("nerf_b200 ops take CUDA tensors (there is no CPU path)")
    if not t.is_contiguous():
        raise NB2Error("tensor must be contiguous")
    return c_vp(t.data_ptr())


def ptr_int(t):
    return 0 if t is None else t.data_ptr()


def f32(t):
    """Contiguous fp32 view/copy of a CUDA tensor."""
    if not t.is_cuda:
        raise NB2Error("nerf_b200 ops take CUDA tensors (there is no CPU path)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def launch_count(device=None):
    return int(load().nb2_launch_count(handle(device)))
