"""The layer-wise engine's Python face: `gemm` binds nb2_gemm_bf16 (include/nerf_b200.h), the generic tcgen05 GEMM from
which every nn.Linear forward / dgrad / wgrad of the training step and of Ref-NeRF is built.  Tensors may be column
views of wider row-major buffers (stride(1) == 1): that is how torch.cat inputs are expressed without copies."""
import ctypes

import torch

from . import _lib
from ._lib import NB2Error, check, handle, load, stream_ptr

ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2


class Program:
    """A launch plan: the GEMMs (and the few other kernels between them) of one pass over a network, recorded once per
    (network, batch size) and replayed with one host call per run of consecutive GEMMs (nb2_gemm_bf16_batch) plus one for
    every split-K reduction (nb2_reduce_splits_batch, deferred to the end: nothing inside a pass reads a weight gradient).

        prog = Program(dev)
        with prog:                        # gemm() / reduce_splits() / call() record instead of launching
            prog.bind(x=x, out=out)       # tensors that differ from run to run; pointers into them are re-based by run()
            ...
        prog.run(x=x2, out=out2)

    Everything else a recorded launch touches (activations, converted weights, workspaces) is kept alive by the plan and
    must not be reallocated by the caller."""

    def __init__(self, dev):
        self.dev = dev
        self.inp = {}
        self._dyn = {}          # key -> (base pointer at record time, bytes)
        self._ops = []          # ["gemm", [descs]] | ["call", fn]
        self._reduces = []
        self._pending = []      # (kind, op index, index inside the op, field, key, byte offset)
        self._patches = []      # (struct inside the final array, field, key, byte offset)
        self._final = None
        self._red = None
        self.keep = []
        self._prev = None

    def __enter__(self):
        global _current
        self._prev, _current = _current, self
        return self

    def __exit__(self, *exc):
        global _current
        _current = self._prev
        if exc[0] is None:
            self._finalize()
        return False

    def bind(self, **tensors):
        for k, t in tensors.items():
            self.inp[k] = t
            self._dyn[k] = (t.data_ptr(), t.numel() * t.element_size())

    def _dynamic(self, ptr):
        if ptr:
            for k, (base, nbytes) in self._dyn.items():
                if base <= ptr < base + nbytes:
                    return k, ptr - base
        return None

    def call(self, fn):
        """fn(): a launch that is not a GEMM; read run-to-run tensors from prog.inp[...] inside it."""
        self._ops.append(["call", fn])

    def _add_gemm(self, d, tensors):
        if not self._ops or self._ops[-1][0] != "gemm":
            self._ops.append(["gemm", []])
        batch = self._ops[-1][1]
        for field in ("out_f32",):
            hit = self._dynamic(getattr(d, field))
            if hit:
                self._pending.append(("gemm", len(self._ops) - 1, len(batch), field, *hit))
        batch.append(d)
        self.keep.extend(t for t in tensors if t is not None)

    def _add_reduce(self, d, tensors):
        hit = self._dynamic(d.out)
        if hit:
            self._pending.append(("reduce", 0, len(self._reduces), "out", *hit))
        self._reduces.append(d)
        self.keep.extend(t for t in tensors if t is not None)

    def _finalize(self):
        final = []
        for kind, payload in self._ops:
            if kind == "gemm":
                arr = (_lib.GemmDesc * len(payload))(*payload)
                final.append((arr, len(payload)))
            else:
                final.append((payload, -1))
        self._final = final
        self._red = (_lib.ReduceDesc * len(self._reduces))(*self._reduces) if self._reduces else None
        for kind, op, idx, field, key, off in self._pending:
            elem = self._red[idx] if kind == "reduce" else final[op][0][idx]
            self._patches.append((elem, field, key, off))
        self.launches = sum(n if n >= 0 else 1 for _, n in final) + len(self._reduces)
        self._ops = self._pending = None

    def run(self, **tensors):
        inp = self.inp
        inp.update(tensors)
        base = {k: inp[k].data_ptr() for k in self._dyn}
        for elem, field, key, off in self._patches:
            setattr(elem, field, base[key] + off)
        lib, h, st = load(), handle(self.dev), stream_ptr(self.dev)
        for a, n in self._final:
            if n >= 0:
                check(lib.nb2_gemm_bf16_batch(h, a, n, st))
            else:
                a()
        if self._red is not None:
            check(lib.nb2_reduce_splits_batch(h, self._red, len(self._red), st))

    def release(self):
        """Drop the run-to-run tensors (the plan itself stays valid)."""
        for k in self.inp:
            self.inp[k] = None


_current = None


def recording():
    return _current


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


def gemm(M, N, segs, bias=None, act=ACT_NONE, mask=None, out_f32=None, out_hi=None, out_lo=None, splits=1, split_stride=0,
         a_rowsum=None, a_rowsum_stride=0):
    """D[M, N] = epi(sum_s A_s B_s^T).  segs: list of (A, a_mn_major, B, b_mn_major, K).  See nb2_gemm_desc.
    a_rowsum (fp32, splits x a_rowsum_stride): per-split sums over K of A's rows (the bias gradient riding on the wgrad)."""
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
    if a_rowsum is not None:
        if a_rowsum.dtype != torch.float32 or not a_rowsum.is_contiguous():
            raise NB2Error("gemm: a_rowsum must be a contiguous fp32 tensor")
        d.a_rowsum_out, d.a_rowsum_stride = a_rowsum.data_ptr(), a_rowsum_stride
    if _current is not None:
        _current._add_gemm(d, [t for A, _, B, _, _ in segs for t in (A, B)] + [bias, mask, out_f32, out_hi, out_lo, a_rowsum])
        return
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
    if _current is not None:
        if accumulate:
            raise NB2Error("reduce_splits(accumulate=True) cannot be recorded: a plan's reductions run as one unordered launch")
        d = _lib.ReduceDesc(ws.data_ptr(), splits, rows, cols, ld_ws, out.stride(0) if out.dim() == 2 else cols, 1 if accumulate else 0,
                            split_stride, _lib.ptr_int(col_perm), out.data_ptr())
        _current._add_reduce(d, [ws, col_perm, out])
        return
    check(load().nb2_reduce_splits(handle(dev), ws.data_ptr(), splits, split_stride, rows, cols, ld_ws, _lib.ptr_int(col_perm),
                                   out.data_ptr(), out.stride(0) if out.dim() == 2 else cols, 1 if accumulate else 0, stream_ptr(dev)))
