"""The layer-wise engine's Python face: `gemm` binds nb2_gemm_bf16 (include/nerf_b200.h), the generic tcgen05 GEMM from
which every nn.Linear forward / dgrad / wgrad of the training step and of Ref-NeRF is built.  Tensors may be column
views of wider row-major buffers (stride(1) == 1): that is how torch.cat inputs are expressed without copies."""
import ctypes

import torch

from . import _lib
from ._lib import NB2Error, check, handle, load, stream_ptr

ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2


def _operand(t, mn_major):
    if t.dtype != torch.bfloat16 or t.dim() != 2 or t.stride(1) != 1:
        raise NB2Error("gemm operands are 2-D bf16 tensors with unit column stride")
    return _lib.GemmOperand(t.data_ptr(), t.stride(0), 1 if mn_major else 0, 0)


def _out(t, dtype):
    if t is None:
        return None, 0
    if t.dtype != dtype or t.dim() != 2 or t.stride(1) != 1:
        raise NB2Error(f"gemm outputs are 2-D {dtype} tensors with unit column stride")
    return t.data_ptr(), t.stride(0)


def gemm(M, N, segs, bias=None, act=ACT_NONE, mask=None, out_f32=None, out_hi=None, out_lo=None, splits=1, split_stride=0):
    """D[M, N] = epi(sum_s A_s B_s^T).  segs: list of (A, a_mn_major, B, b_mn_major, K).  See nb2_gemm_desc."""
    if not 1 <= len(segs) <= _lib.GEMM_MAX_SEG:
        raise NB2Error(f"gemm: 1..{_lib.GEMM_MAX_SEG} segments")
    dev = segs[0][0].device
    d = _lib.GemmDesc()
    d.M, d.N, d.n_seg = M, N, len(segs)
    for i, (A, a_mn, B, b_mn, K) in enumerate(segs):
        d.seg[i].a, d.seg[i].b, d.seg[i].K = _operand(A, a_mn), _operand(B, b_mn), K
    if bias is not None:
        if bias.dtype != torch.float32 or not bias.is_contiguous():
            raise NB2Error("gemm: bias must be a contiguous fp32 vector")
        d.bias = bias.data_ptr()
    d.act = act
    if mask is not None:
        d.mask, d.ld_mask = _out(mask, torch.bfloat16)
    d.out_f32, d.ld_f32 = _out(out_f32, torch.float32)
    d.out_hi, d.ld_16 = _out(out_hi, torch.bfloat16)
    if out_lo is not None:
        lo_ptr, lo_ld = _out(out_lo, torch.bfloat16)
        if lo_ld != d.ld_16:
            raise NB2Error("gemm: out_hi and out_lo must share their row stride")
        d.out_lo = lo_ptr
    d.splits, d.split_stride = splits, split_stride
    check(load().nb2_gemm_bf16(handle(dev), ctypes.byref(d), stream_ptr(dev)))


def to_bf16(src, ld_dst=None, want_lo=True, col_perm=None, out=None):
    """fp32 (rows, cols) [row stride may exceed cols] -> (hi, lo) bf16 (rows, ld_dst), pad columns zero."""
    if src.dtype != torch.float32 or src.dim() != 2 or src.stride(1) != 1:
        raise NB2Error("to_bf16: 2-D fp32 tensor with unit column stride")
    rows, cols = src.shape
    ld_dst = (cols + 7) // 8 * 8 if ld_dst is None else ld_dst
    dev = src.device
    if out is None:
        alloc = torch.zeros if col_perm is not None else torch.empty
        hi = alloc((rows, ld_dst), dtype=torch.bfloat16, device=dev)
        lo = alloc((rows, ld_dst), dtype=torch.bfloat16, device=dev) if want_lo else None
    else:
        hi, lo = out
    check(load().nb2_to_bf16(handle(dev), src.data_ptr(), rows, cols, src.stride(0), _lib.ptr_int(col_perm), hi.data_ptr(),
                             _lib.ptr_int(lo), ld_dst, stream_ptr(dev)))
    return hi, lo


def reduce_splits(ws, splits, split_stride, rows, cols, ld_ws, out, col_perm=None, accumulate=False):
    dev = ws.device
    check(load().nb2_reduce_splits(handle(dev), ws.data_ptr(), splits, split_stride, rows, cols, ld_ws, _lib.ptr_int(col_perm),
                                   out.data_ptr(), out.stride(0) if out.dim() == 2 else cols, 1 if accumulate else 0, stream_ptr(dev)))
