"""The differentiable, layer-wise engine behind MipNeRF.forward / ProposalNetwork.forward when gradients are wanted
(the `run()` closure of the reference's train.py:164-199 and its `loss.backward()`).

The fused inference kernels keep activations in shared / tensor memory and record nothing; training needs every layer's
input for the weight gradients.  Here each nn.Linear is one launch of the generic tcgen05 GEMM (nb2_gemm_bf16,
csrc/nb2_gemm.cu) whose epilogue writes the activation as bf16 hi (+ lo residual) rows -- exactly the operand format of the
next layer's forward, of the dgrad and of the wgrad -- so the forward pass saves what the backward pass reads and nothing is
converted or transposed in between:

    forward   H_{l+1} = relu(H_l W_l^T + b_l)                 A = H_l (K-major),   B = W_l (K-major)
    dgrad     dH_l    = (dH_{l+1} W_l) * (H_l > 0)            A = dH_{l+1},        B = W_l (MN-major)   relu mask = saved H_l
    wgrad     dW_l    = dH_{l+1}^T H_l                        A = dH_{l+1} (MN),   B = H_l (MN), split-K + deterministic reduce
    bgrad     db_l    = dH_{l+1}^T 1                          the wgrad GEMM against a column of ones

torch.cat inputs of the reference (mip_model.py:55 skip connection, :59 bottleneck + encoded direction) are column ranges
of one wider buffer; the matching weight columns are permuted once when the weights are converted, and the gradient is
permuted back by the split-K reduction.  precision 'bf16x3' (default): every product as lo*hi + hi*lo + hi*hi (three K
segments, 16 significant bits, no fp16 range problem for gradients); 'bf16': one pass (the analogue of the reference's
autocast training, train.py:201-207, without needing a GradScaler because bf16 keeps fp32's exponent).
"""
import ctypes

import torch

from . import _lib, linear
from ._lib import check, handle, load, stream_ptr

BF16 = torch.bfloat16


def _pad8(n):
    return (n + 7) // 8 * 8


def _empty16(rows, cols, dev, lo=True):
    return (torch.empty((rows, cols), dtype=BF16, device=dev), torch.empty((rows, cols), dtype=BF16, device=dev) if lo else None)


def encode(x, x_col0, levels, normalize, hi, lo):
    """Rows [x, sin(2^l x), cos(2^l x) ...] of points x[:, x_col0:x_col0+3] into the (column view) hi / lo."""
    dev = x.device
    check(load().nb2_encode_bf16(handle(dev), x.data_ptr(), x.stride(0), x_col0, x.shape[0], levels, 1 if normalize else 0, hi.data_ptr(),
                                 _lib.ptr_int(lo), hi.stride(0), hi.shape[1], stream_ptr(dev)))


class PackedLinear:
    """bf16 hi / lo image of one nn.Linear in the engine's column order, refreshed when the parameter changes."""

    def __init__(self, lin, col_perm=None, in_pad=None):
        self.lin = lin
        self.out_f, self.in_f = lin.weight.shape
        self.in_pad = _pad8(self.in_f) if in_pad is None else in_pad
        self.perm_host = col_perm
        self.perm = None
        self.key = None
        self.hi = self.lo = None

    def sync(self):
        w = self.lin.weight
        key = (w.data_ptr(), w._version, w.device)
        if key != self.key:
            if self.perm_host is not None and (self.perm is None or self.perm.device != w.device):
                self.perm = self.perm_host.to(device=w.device, dtype=torch.int32)
            same = self.hi is not None and self.hi.device == w.device
            # converted in place: launch plans hold pointers to these buffers
            self.hi, self.lo = linear.to_bf16(w.detach(), ld_dst=self.in_pad, col_perm=self.perm, out=(self.hi, self.lo) if same else None)
            self.key = key
        return self

    def cols(self, c0, c1):
        return self.hi[:, c0:c1], self.lo[:, c0:c1]


def sync_all(packed):
    """Refresh every stale weight image of `packed` (PackedLinear list): images that already own their buffers are converted
    by ONE launch (nb2_to_bf16_batch) -- after an optimizer step all of a network's matrices are stale at once -- the others
    (first use, device change) through PackedLinear.sync()."""
    stale = []
    for p in packed:
        if type(p) is not PackedLinear:      # packed heads (ref_model._PackedCat) concatenate several matrices first
            p.sync()
            continue
        w = p.lin.weight
        key = (w.data_ptr(), w._version, w.device)
        if key == p.key:
            continue
        if p.hi is None or p.hi.device != w.device or (p.perm_host is not None and (p.perm is None or p.perm.device != w.device)):
            p.sync()
            continue
        stale.append((p, w, key))
    by_dev = {}
    for item in stale:
        by_dev.setdefault(item[1].device, []).append(item)
    for dev, items in by_dev.items():
        arr = (_lib.ToBf16Desc * len(items))()
        for d, (p, w, _) in zip(arr, items):
            wd = w.detach()
            if wd.dtype != torch.float32 or wd.dim() != 2 or wd.stride(1) != 1:
                raise _lib.NB2Error("sync_all: weights are 2-D fp32 tensors with unit column stride")
            d.src, d.rows, d.ld_src, d.cols, d.ld_dst = wd.data_ptr(), wd.shape[0], wd.stride(0), wd.shape[1], p.in_pad
            d.col_perm, d.hi, d.lo = _lib.ptr_int(p.perm), p.hi.data_ptr(), p.lo.data_ptr()
        check(load().nb2_to_bf16_batch(handle(dev), arr, len(items), stream_ptr(dev)))
        for p, _, key in items:
            p.key = key


def _fwd_segs(x, w, K, x3):
    """x, w: (hi, lo) pairs; small terms first."""
    if x3:
        return [(x[1], False, w[0], False, K), (x[0], False, w[1], False, K), (x[0], False, w[0], False, K)]
    return [(x[0], False, w[0], False, K)]


def _dgrad_segs(dy, w, K, x3):
    if x3:
        return [(dy[1], False, w[0], True, K), (dy[0], False, w[1], True, K), (dy[0], False, w[0], True, K)]
    return [(dy[0], False, w[0], True, K)]


class _Workspace:
    """Split-K partial sums of the weight gradients (one buffer per device, grown on demand)."""
    bufs = {}

    @classmethod
    def get(cls, dev, floats):
        if linear.recording() is not None:       # a launch plan defers its reductions: every wgrad keeps its own partial sums
            return torch.empty(floats, dtype=torch.float32, device=dev)
        b = cls.bufs.get(dev)
        if b is None or b.numel() < floats:
            b = torch.empty(floats, dtype=torch.float32, device=dev)
            cls.bufs[dev] = b
        return b


class _Ones:
    """(rows, 8) bf16 with column 0 = 1: the B operand that turns the wgrad GEMM into the bias gradient (column sums of dY)."""
    bufs = {}

    @classmethod
    def get(cls, dev, rows):
        b = cls.bufs.get(dev)
        if b is None or b.shape[0] < rows:
            b = torch.zeros((rows, 8), dtype=BF16, device=dev)
            b[:, 0] = 1.0
            cls.bufs[dev] = b
        return b[:rows]


def _row_targets(grad, n_out):
    """grad: one tensor for all n_out rows, or [(row0, row1, tensor)] when the rows belong to several parameters (packed heads)."""
    return grad if isinstance(grad, list) else [(0, n_out, grad)]


def bgrad(dy, n_out, rows, x3, grad_b, sm_count=148):
    """grad_b (n_out,) fp32 = column sums of dy over `rows` samples, as dy^T @ ones on the tensor cores (split-K, deterministic)."""
    dev = dy[0].device
    ones = _Ones.get(dev, rows)
    m_pad = (n_out + 127) // 128 * 128
    splits = max(1, min(sm_count // (m_pad // 128), (rows + 255) // 256))
    ws = _Workspace.get(dev, splits * m_pad * 32)
    out = ws[: splits * m_pad * 32].view(splits * m_pad, 32)
    segs = [(dy[0], True, ones, True, rows)] + ([(dy[1], True, ones, True, rows)] if x3 else [])
    linear.gemm(n_out, 8, segs, out_f32=out[:n_out], splits=splits, split_stride=m_pad * 32)
    for r0, r1, g in _row_targets(grad_b, n_out):
        linear.reduce_splits(ws[r0 * 32:], splits, m_pad * 32, r1 - r0, 1, 32, g.view(r1 - r0, 1))


def wgrad(dy, x, n_out, n_in, rows, x3, grad_w, perm=None, sm_count=148, grad_b=None):
    """grad_w (n_out, n_in) fp32 = dy^T x over `rows` samples; dy (rows, >= n_out), x (rows, ld >= n_in) as (hi, lo).
    grad_b (n_out,): the bias gradient dy^T 1 of the same layer.  In the split precision it rides on this GEMM (one extra
    N = 16 MMA per k-step against a tile of ones, nb2_gemm_desc.a_rowsum_out) instead of re-reading dy in a GEMM of its own."""
    dev = x[0].device
    ld_ws = (x[0].shape[1] + 31) // 32 * 32
    n_cols = x[0].shape[1]
    m_pad = (n_out + 127) // 128 * 128
    tiles = (m_pad // 128) * ((ld_ws + 255) // 256)
    splits = max(1, min(sm_count // tiles, (rows + 255) // 256))
    fuse_b = grad_b is not None and x3
    n_ws = splits * m_pad * ld_ws
    ws = _Workspace.get(dev, n_ws + (splits * m_pad if fuse_b else 0))
    out = ws[:n_ws].view(splits * m_pad, ld_ws)
    rs = ws[n_ws:n_ws + splits * m_pad] if fuse_b else None
    if x3:
        segs = [(dy[1], True, x[0], True, rows), (dy[0], True, x[1], True, rows), (dy[0], True, x[0], True, rows)]
    else:
        segs = [(dy[0], True, x[0], True, rows)]
    linear.gemm(n_out, n_cols, segs, out_f32=out[:n_out], splits=max(splits, 1) if splits > 1 else 1, split_stride=m_pad * ld_ws,
                a_rowsum=rs, a_rowsum_stride=m_pad)
    for r0, r1, g in _row_targets(grad_w, n_out):
        linear.reduce_splits(ws[r0 * ld_ws:], splits, m_pad * ld_ws, r1 - r0, n_in, ld_ws, g, col_perm=perm)
    if fuse_b:
        for r0, r1, g in _row_targets(grad_b, n_out):
            linear.reduce_splits(rs[r0:], splits, m_pad, r1 - r0, 1, 1, g.view(r1 - r0, 1))
    elif grad_b is not None:
        bgrad(dy, n_out, rows, x3, grad_b, sm_count=sm_count)


class _Plan:
    """The recorded forward and backward launch plans of one network at one batch size, with the activations between them."""

    def __init__(self):
        self.fwd = self.bwd = None
        self.acts = None
        self.out = None
        self.busy = False      # between a forward and its (first) backward the activations belong to that autograd node
        self.epoch = 0         # forward passes run on this plan: a later backward of an older pass finds its activations gone
        self.want_dx = False   # the backward plan also produces d loss / d positions
        self.pts = None


class _Engine:
    """Shared plan bookkeeping.  Subclasses record `_forward(prog, plan, n, x3)` and `_backward(prog, plan, n, x3, flat)`."""

    keep_last_acts = False     # tests: keep a reference to the saved activations of the last forward (relu patterns)
    last_acts = None
    max_plans = 2              # batch sizes kept recorded per network (activations stay allocated with their plan)

    def _init_plans(self):
        self.plans = {}
        self.grad_shapes = [tuple(p.shape) for l in self.lins for p in (l.weight, l.bias)]
        self.grad_numel = sum(l.weight.numel() + l.bias.numel() for l in self.lins)

    def clear_plans(self):
        self.plans.clear()

    def _plan(self, n, x3, dev, want_dx=False):
        key = (n, x3, dev, want_dx) + tuple(p.data_ptr() for l in self.lins for p in (l.weight, l.bias))
        plan = self.plans.get(key)
        if plan is None or plan.busy:
            # busy: a second forward before the first one's backward (or a forward whose backward never ran) records anew
            self.plans.pop(key, None)
            while len(self.plans) >= self.max_plans:
                self.plans.pop(next(iter(self.plans)))
            plan = self.plans[key] = _Plan()
            plan.want_dx = want_dx
        return plan

    def forward(self, pts, x3, want_dx=False):
        """want_dx: the backward pass also produces the gradient of the positions (pts columns 0..2)."""
        dev, n = pts.device, pts.shape[0]
        pts = pts.contiguous()
        plan = self._plan(n, x3, dev, want_dx)
        sync_all(self.packed)
        out = torch.empty((n, self.out_cols), dtype=torch.float32, device=dev)
        if plan.fwd is None:
            with linear.Program(dev) as prog:
                prog.bind(pts=pts, out=out)
                plan.acts = self._forward(prog, n, x3, out)
            plan.fwd = prog
        plan.fwd.run(pts=pts, out=out)
        plan.fwd.release()
        plan.out, plan.busy, plan.pts = out, True, (pts if want_dx else None)
        plan.epoch += 1
        return out, plan

    def backward(self, plan, g_out, x3):
        """g_out: gradient of forward's output -> list of (grad_weight, grad_bias) in state_dict order."""
        dev, n = g_out.device, g_out.shape[0]
        g_out = g_out.contiguous()
        flat = torch.empty(self.grad_numel, dtype=torch.float32, device=dev)
        dyn = dict(g=g_out, out=plan.out, grads=flat)
        d_pts = None
        if plan.want_dx:
            d_pts = torch.empty((n, 3), dtype=torch.float32, device=dev)
            dyn.update(pts=plan.pts, d_pts=d_pts)
        if plan.bwd is None:
            with linear.Program(dev) as prog:
                prog.bind(**dyn)
                d_enc = self._backward(prog, plan.acts, n, x3, self._grad_views(flat), plan.want_dx)
                if plan.want_dx:
                    levels, ld = self.pos_levels, d_enc.shape[1]
                    prog.call(lambda: check(load().nb2_encode_backward(handle(dev), prog.inp["pts"].data_ptr(), prog.inp["pts"].stride(0), 0, n,
                                                                       levels, d_enc.data_ptr(), ld, prog.inp["d_pts"].data_ptr(), stream_ptr(dev))))
            plan.bwd = prog
        plan.bwd.run(**dyn)
        plan.bwd.release()
        plan.busy = False      # the next forward may take the plan; until then this pass can be differentiated again
        return self._grad_views(flat), d_pts

    def _grad_views(self, flat):
        views, off = [], 0
        for l in self.lins:
            nw, nb = l.weight.numel(), l.bias.numel()
            views.append((flat[off:off + nw].view(l.weight.shape), flat[off + nw:off + nw + nb]))
            off += nw + nb
        return views


class ProposalEngine(_Engine):
    """ProposalNetwork(10, 256): 63 -> 256 -> 256 -> 256 -> 256 -> 1 (nerf/addtional.py:61-72,88-96)."""

    out_cols = 1

    def __init__(self, module):
        self.m = module
        L = module.layers
        self.lins = [L[0], L[2], L[4], L[6], L[8]]
        self.packed = [PackedLinear(l) for l in self.lins]
        self.levels = self.pos_levels = module.position_flevel
        self._init_plans()

    def _forward(self, prog, n, x3, sigma):
        """pts (n, 3) fp32 -> sigma (n, 1) fp32; returns the saved activations."""
        dev = prog.dev
        W = self.packed
        H = self.lins[0].out_features
        enc_w = _pad8(3 + 6 * self.levels)
        E = _empty16(n, enc_w, dev, x3)
        prog.call(lambda: encode(prog.inp["pts"], 0, self.levels, False, E[0], E[1]))
        acts = [E]
        x, K = E, 3 + 6 * self.levels
        for l in range(4):
            y = _empty16(n, H, dev, x3)
            linear.gemm(n, H, _fwd_segs(x, (W[l].hi, W[l].lo), K, x3), bias=self.lins[l].bias.detach(), act=linear.ACT_RELU,
                        out_hi=y[0], out_lo=y[1])
            acts.append(y)
            x, K = y, H
        linear.gemm(n, 1, _fwd_segs(x, (W[4].hi, W[4].lo), H, x3), bias=self.lins[4].bias.detach(), out_f32=sigma)
        return acts

    def _backward(self, prog, acts, n, x3, grads, want_dx=False):
        dev = prog.dev
        W = self.packed
        H = self.lins[0].out_features
        sm = torch.cuda.get_device_properties(dev).multi_processor_count
        ds = _empty16(n, 8, dev, x3)
        prog.call(lambda: linear.to_bf16(prog.inp["g"].view(n, 1), ld_dst=8, want_lo=x3, out=ds))
        wgrad(ds, acts[4], 1, H, n, x3, grads[4][0], sm_count=sm, grad_b=grads[4][1])
        dy = _empty16(n, H, dev, x3)
        linear.gemm(n, H, _dgrad_segs(ds, (W[4].hi, W[4].lo), 1, x3), mask=acts[4][0], out_hi=dy[0], out_lo=dy[1])
        for l in (3, 2, 1, 0):
            x = acts[l]
            wgrad(dy, x, H, self.lins[l].in_features, n, x3, grads[l][0], sm_count=sm, grad_b=grads[l][1])
            if l > 0:
                dx = _empty16(n, H, dev, x3)
                linear.gemm(n, H, _dgrad_segs(dy, (W[l].hi, W[l].lo), H, x3), mask=x[0], out_hi=dx[0], out_lo=dx[1])
                dy = dx
        if not want_dx:
            return None
        enc_w = _pad8(3 + 6 * self.levels)
        d_enc = torch.empty((n, enc_w), dtype=torch.float32, device=dev)       # gradient of the encoded positions
        linear.gemm(n, enc_w, _dgrad_segs(dy, (W[0].hi, W[0].lo), H, x3), out_f32=d_enc)
        return d_enc


class NerfEngine(_Engine):
    """MipNeRF(10, 4, 256): the vanilla 8x256 NeRF MLP with skip connection and heads (nerf/mip_model.py:14-60)."""

    out_cols = 4

    def __init__(self, module):
        self.m = module
        b1, b2 = module.lin_block1, module.lin_block2
        self.lins = [b1[0], b1[2], b1[4], b1[6], b2[0], b2[2], b2[4], module.bottle_neck[0], module.opacity_head[0],
                     module.rgb_layer[0], module.rgb_layer[2]]
        self.pl, self.dl = module.position_flevel, module.direction_flevel
        self.pos_levels = self.pl
        self.enc = 3 + 6 * self.pl
        self.denc = 3 + 6 * self.dl
        H = self.lins[0].out_features
        self.H = H
        # lin_block2.0 multiplies cat(enc, h) (mip_model.py:55); the engine's buffer is [h | enc | pad]
        perm4 = torch.cat((torch.arange(self.enc) + H, torch.arange(H)))
        self.packed = [PackedLinear(l) for l in self.lins]
        self.packed[4] = PackedLinear(self.lins[4], col_perm=perm4, in_pad=H + _pad8(self.enc))
        self._init_plans()

    def _forward(self, prog, n, x3, out):
        """pts (n, 6) fp32 = [xyz, dir] -> out (n, 4) fp32 = [sigmoid rgb, raw sigma]; returns the saved activations."""
        dev = prog.dev
        W = self.packed
        H, enc_w, denc_w = self.H, _pad8(self.enc), _pad8(self.denc)
        b = [l.bias.detach() for l in self.lins]
        C5 = _empty16(n, H + enc_w, dev, x3)                      # [h4 | enc]
        E = (C5[0][:, H:], C5[1][:, H:] if x3 else None)
        C9 = _empty16(n, H + denc_w, dev, x3)                     # [bottleneck | enc(dir)]
        D = (C9[0][:, H:], C9[1][:, H:] if x3 else None)

        def encodings():
            pts = prog.inp["pts"]
            encode(pts, 0, self.pl, False, E[0], E[1])
            encode(pts, 3, self.dl, True, D[0], D[1])

        prog.call(encodings)
        wl = lambda i: (W[i].hi, W[i].lo)
        h1, h2, h3 = (_empty16(n, H, dev, x3) for _ in range(3))
        linear.gemm(n, H, _fwd_segs(E, wl(0), self.enc, x3), bias=b[0], act=linear.ACT_RELU, out_hi=h1[0], out_lo=h1[1])
        linear.gemm(n, H, _fwd_segs(h1, wl(1), H, x3), bias=b[1], act=linear.ACT_RELU, out_hi=h2[0], out_lo=h2[1])
        linear.gemm(n, H, _fwd_segs(h2, wl(2), H, x3), bias=b[2], act=linear.ACT_RELU, out_hi=h3[0], out_lo=h3[1])
        h4 = (C5[0][:, :H], C5[1][:, :H] if x3 else None)
        linear.gemm(n, H, _fwd_segs(h3, wl(3), H, x3), bias=b[3], act=linear.ACT_RELU, out_hi=h4[0], out_lo=h4[1])
        h5, h6, h7 = (_empty16(n, H, dev, x3) for _ in range(3))
        linear.gemm(n, H, _fwd_segs(C5, wl(4), H + self.enc, x3), bias=b[4], act=linear.ACT_RELU, out_hi=h5[0], out_lo=h5[1])
        linear.gemm(n, H, _fwd_segs(h5, wl(5), H, x3), bias=b[5], act=linear.ACT_RELU, out_hi=h6[0], out_lo=h6[1])
        linear.gemm(n, 256, _fwd_segs(h6, wl(6), H, x3), bias=b[6], act=linear.ACT_RELU, out_hi=h7[0], out_lo=h7[1])
        linear.gemm(n, 1, _fwd_segs(h7, wl(8), 256, x3), bias=b[8], out_f32=out[:, 3:])                     # opacity_head
        bn = (C9[0][:, :256], C9[1][:, :256] if x3 else None)
        linear.gemm(n, 256, _fwd_segs(h7, wl(7), 256, x3), bias=b[7], out_hi=bn[0], out_lo=bn[1])           # bottle_neck (linear)
        t = _empty16(n, 128, dev, x3)
        linear.gemm(n, 128, _fwd_segs(C9, wl(9), 256 + self.denc, x3), bias=b[9], act=linear.ACT_RELU, out_hi=t[0], out_lo=t[1])
        linear.gemm(n, 3, _fwd_segs(t, wl(10), 128, x3), bias=b[10], act=linear.ACT_SIGMOID, out_f32=out[:, :3])
        return dict(E=E, h1=h1, h2=h2, h3=h3, C5=C5, h5=h5, h6=h6, h7=h7, C9=C9, t=t)

    def _backward(self, prog, a, n, x3, grads, want_dx=False):
        dev = prog.dev
        W = self.packed
        H = self.H
        sm = torch.cuda.get_device_properties(dev).multi_processor_count
        wl = lambda i: (W[i].hi, W[i].lo)

        def wg(i, dy, x, n_out, n_in, perm=None):
            wgrad(dy, x, n_out, n_in, n, x3, grads[i][0], perm=perm, sm_count=sm, grad_b=grads[i][1])

        dz, ds = _empty16(n, 8, dev, x3), _empty16(n, 8, dev, x3)
        prog.call(lambda: check(load().nb2_nerf_head_backward(handle(dev), prog.inp["out"].data_ptr(), prog.inp["g"].data_ptr(), n,
                                                              dz[0].data_ptr(), _lib.ptr_int(dz[1]), ds[0].data_ptr(), _lib.ptr_int(ds[1]),
                                                              stream_ptr(dev))))
        wg(10, dz, a["t"], 3, 128)                                                                   # rgb_layer.2
        dt = _empty16(n, 128, dev, x3)
        linear.gemm(n, 128, _dgrad_segs(dz, wl(10), 3, x3), mask=a["t"][0], out_hi=dt[0], out_lo=dt[1])
        wg(9, dt, a["C9"], 128, 256 + self.denc)                                                     # rgb_layer.0
        db = _empty16(n, 256, dev, x3)
        w9 = W[9].cols(0, 256)
        linear.gemm(n, 256, _dgrad_segs(dt, w9, 128, x3), out_hi=db[0], out_lo=db[1])                # -> bottleneck (no activation)
        wg(7, db, a["h7"], 256, 256)                                                                 # bottle_neck
        wg(8, ds, a["h7"], 1, 256)                                                                   # opacity_head
        d7 = _empty16(n, 256, dev, x3)
        linear.gemm(n, 256, _dgrad_segs(db, wl(7), 256, x3) + _dgrad_segs(ds, wl(8), 1, x3), mask=a["h7"][0], out_hi=d7[0], out_lo=d7[1])
        wg(6, d7, a["h6"], 256, H)
        d6 = _empty16(n, H, dev, x3)
        linear.gemm(n, H, _dgrad_segs(d7, wl(6), 256, x3), mask=a["h6"][0], out_hi=d6[0], out_lo=d6[1])
        wg(5, d6, a["h5"], H, H)
        d5 = _empty16(n, H, dev, x3)
        linear.gemm(n, H, _dgrad_segs(d6, wl(5), H, x3), mask=a["h5"][0], out_hi=d5[0], out_lo=d5[1])
        wg(4, d5, a["C5"], H, H + self.enc, perm=W[4].perm)                                          # lin_block2.0 on cat(enc, h)
        d4 = _empty16(n, H, dev, x3)
        linear.gemm(n, H, _dgrad_segs(d5, W[4].cols(0, H), H, x3), mask=a["C5"][0][:, :H], out_hi=d4[0], out_lo=d4[1])
        wg(3, d4, a["h3"], H, H)
        d3 = _empty16(n, H, dev, x3)
        linear.gemm(n, H, _dgrad_segs(d4, wl(3), H, x3), mask=a["h3"][0], out_hi=d3[0], out_lo=d3[1])
        wg(2, d3, a["h2"], H, H)
        d2 = _empty16(n, H, dev, x3)
        linear.gemm(n, H, _dgrad_segs(d3, wl(2), H, x3), mask=a["h2"][0], out_hi=d2[0], out_lo=d2[1])
        wg(1, d2, a["h1"], H, H)
        d1 = _empty16(n, H, dev, x3)
        linear.gemm(n, H, _dgrad_segs(d2, wl(1), H, x3), mask=a["h1"][0], out_hi=d1[0], out_lo=d1[1])
        wg(0, d1, a["E"], H, self.enc)
        if not want_dx:
            return None
        enc_w = _pad8(self.enc)                                                    # the encoding enters lin_block1.0 and the skip layer
        d_enc = torch.empty((n, enc_w), dtype=torch.float32, device=dev)
        linear.gemm(n, enc_w, _dgrad_segs(d1, wl(0), H, x3) + _dgrad_segs(d5, W[4].cols(H, H + enc_w), H, x3), out_f32=d_enc)
        return d_enc


class _MLPFunction(torch.autograd.Function):
    """forward(engine, x3, pts, *params): params are passed so autograd routes their gradients; values come from the module.
    When pts requires a gradient, backward also returns d loss / d positions (RefNeRF.get_grad on the proposal network,
    train.py:165-168)."""

    @staticmethod
    def forward(ctx, engine, x3, pts, *params):
        out, plan = engine.forward(pts, x3, want_dx=ctx.needs_input_grad[2])
        ctx.engine, ctx.x3, ctx.plan, ctx.epoch, ctx.pts_cols = engine, x3, plan, plan.epoch, pts.shape[1]
        if engine.keep_last_acts:
            engine.last_acts = plan.acts
        return out

    @staticmethod
    def backward(ctx, g):
        if ctx.plan.epoch != ctx.epoch:
            raise _lib.NB2Error("the layer-wise engine keeps one set of activations per recorded plan: this network ran forward again "
                                "at this batch size after this pass had been differentiated, so its activations are gone")
        grads, d_pts = ctx.engine.backward(ctx.plan, g, ctx.x3)
        flat = []
        for gw, gb in grads:
            flat += [gw, gb]
        if d_pts is not None and ctx.pts_cols != 3:        # [xyz, dir] rows: no gradient is produced for the direction columns
            full = torch.zeros((d_pts.shape[0], ctx.pts_cols), dtype=torch.float32, device=d_pts.device)
            full[:, :3] = d_pts
            d_pts = full
        return (None, None, d_pts, *flat)


def train_engine_of(module, engine_cls):
    eng = module.__dict__.get("_nb2_train_engine")
    if eng is None:
        eng = engine_cls(module)
        module.__dict__["_nb2_train_engine"] = eng
    return eng


def differentiable_forward(module, engine_cls, pts2d, precision):
    eng = train_engine_of(module, engine_cls)
    if precision not in (None, "bf16x3", "bf16"):
        raise _lib.NB2Error(f"the differentiable layer-wise engine runs 'bf16x3' (default) or 'bf16', not {precision!r}")
    params = [p for l in eng.lins for p in (l.weight, l.bias)]
    return _MLPFunction.apply(eng, precision != "bf16", pts2d, *params)


# ---- autograd wrappers of the ray ops (reference-named entry points call these when a gradient is required) ----------
class WeightsFromSigma(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sigma, z, dirs, act):
        from . import ops
        w = ops.weights_from_sigma(sigma, z, dirs, act)
        ctx.save_for_backward(sigma, z, dirs if dirs is not None else torch.empty(0, device=z.device))
        ctx.act, ctx.has_dirs = act, dirs is not None
        return w

    @staticmethod
    def backward(ctx, g):
        from . import ops
        sigma, z, dirs = ctx.saved_tensors
        sigma, z, g = _lib.f32(sigma), _lib.f32(z), _lib.f32(g)
        dirs = _lib.f32(dirs) if ctx.has_dirs else None
        R, P = z.shape
        d = torch.empty_like(sigma)
        check(load().nb2_weights_from_sigma_backward(handle(z.device), sigma.data_ptr(), z.data_ptr(), _lib.ptr_int(dirs),
                                                     dirs.shape[-1] if dirs is not None else 0, R, P, ops._ACTS[ctx.act], g.data_ptr(),
                                                     d.data_ptr(), stream_ptr(z.device)))
        return d, None, None, None


class Composite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgbo, z, dirs, white_bkg):
        from . import ops
        rgb, w, _, _ = ops.composite(rgbo, z, dirs, white_bkg=white_bkg)
        ctx.save_for_backward(rgbo, z, dirs)
        ctx.white = white_bkg
        return rgb, w

    @staticmethod
    def backward(ctx, g_rgb, g_w):
        rgbo, z, dirs = ctx.saved_tensors
        rgbo, z, dirs = _lib.f32(rgbo), _lib.f32(z), _lib.f32(dirs)
        R, P = z.shape
        g_rgb = _lib.f32(g_rgb) if g_rgb is not None else torch.zeros((R, 3), dtype=torch.float32, device=z.device)
        g_w = _lib.f32(g_w) if g_w is not None else None
        d = torch.empty_like(rgbo)
        check(load().nb2_composite_backward(handle(z.device), rgbo.data_ptr(), z.data_ptr(), dirs.data_ptr(), dirs.shape[-1], R, P,
                                            _lib.WHITE_BKG if ctx.white else 0, g_rgb.data_ptr(), _lib.ptr_int(g_w), d.data_ptr(),
                                            stream_ptr(z.device)))
        return d, None, None, None


class MaxBlur(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w, alpha):
        from . import ops
        ctx.save_for_backward(w)
        return ops.max_blur(w, alpha)

    @staticmethod
    def backward(ctx, g):
        (w,) = ctx.saved_tensors
        w, g = _lib.f32(w), _lib.f32(g)
        P = w.shape[-1]
        d = torch.empty_like(w)
        check(load().nb2_max_blur_backward(handle(w.device), w.data_ptr(), g.data_ptr(), w.numel() // P, P, d.data_ptr(), stream_ptr(w.device)))
        return d, None


class GetBounds(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w, inds):
        from . import ops
        ctx.save_for_backward(inds)
        ctx.P = w.shape[-1]
        return ops.get_bounds(w, inds)

    @staticmethod
    def backward(ctx, g):
        (inds,) = ctx.saved_tensors
        inds = inds.to(torch.int64).contiguous()
        g = _lib.f32(g)
        R, K = inds.shape
        d = torch.empty((R, ctx.P), dtype=torch.float32, device=g.device)
        check(load().nb2_get_bounds_backward(handle(g.device), inds.data_ptr(), g.data_ptr(), R, ctx.P, K, d.data_ptr(), stream_ptr(g.device)))
        return d, None


def allreduce_gradients(modules, group=None, average=True):
    """ONE flat all-reduce of every gradient of `modules` (the collective DistributedDataParallel performs for the reference,
    ddp_train.py:98; 530,052 fp32 values = 2.12 MB for MipNeRF).  NCCL over NVLink on GPUs, gloo in the CPU tests."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    params = [p for m in modules for p in m.parameters() if p.grad is not None]
    if not params:
        return 0
    # The engine's backward hands autograd VIEWS of one flat gradient buffer per network; when the optimizer's gradients still
    # are those views (zero_grad(set_to_none=True), the default), consecutive gradients are adjacent in memory and every such
    # run is reduced in place: no gather / scatter copies around the collective.  Anything else goes through one packed copy.
    world = dist.get_world_size(group)
    runs, loose = [], []
    for p in params:
        g = p.grad
        if not (g.is_contiguous() and g.dtype == torch.float32):
            loose.append(p)
            continue
        last = runs[-1] if runs else None
        if (last is not None and last["storage"] == g.untyped_storage().data_ptr() and last["device"] == g.device
                and last["end"] == g.storage_offset()):
            last["end"] += g.numel()
        else:
            runs.append({"storage": g.untyped_storage().data_ptr(), "device": g.device, "first": g, "start": g.storage_offset(),
                         "end": g.storage_offset() + g.numel()})
    total = 0
    for r in runs:
        n = r["end"] - r["start"]
        flat = torch.empty(0, dtype=torch.float32, device=r["device"]).set_(r["first"].untyped_storage(), r["start"], (n,))
        dist.all_reduce(flat, group=group)
        if average:
            flat /= world
        total += n
    if loose:
        flat = torch.cat([p.grad.reshape(-1).float() for p in loose])
        dist.all_reduce(flat, group=group)
        if average:
            flat /= world
        off = 0
        for p in loose:
            n = p.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
        total += flat.numel()
    return total
