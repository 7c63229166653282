"""Mirror of the reference's nerf/ref_model.py: Ref-NeRF (forward; SURVEY 8f-3, BASELINE configs[3]).

Same constructor, submodule names and state_dict keys as the reference (nerf/ref_model.py:16-65).  forward() runs on
the layer-wise engine: every nn.Linear is one launch of the generic tcgen05 GEMM (nb2_gemm_bf16, bf16 hi + lo operands,
fp32 accumulation), the integrated directional encoding is `ide_kernel`, and the elementwise steps between the MLPs
(normal / reflection / roughness, colour composition) are three small CUDA kernels (csrc/nb2_refnerf.cu).  The
reference's torch.cat inputs are column ranges of wider buffers (the matching weight columns are permuted once).

Training (train.py:164-199 with is_ref_model): forward() under autograd is one torch.autograd.Function over the same
recorded forward plan; its backward is a second recorded plan on the same engine -- dgrad / wgrad / bias-gradient GEMMs for
the 19 linear layers, and three CUDA kernels for the glue (colour composition, normal / reflection / integrated directional
encoding / roughness, positional encoding) -- and returns the gradient of every parameter AND of the sample positions, which
is what the reference's RefNeRF.get_grad (torch.autograd.grad(density, positions), ref_model.py:118-124: first order, the
result carries no graph) and its normal / back-face losses need.  Gradients w.r.t. the view directions are not produced
(the reference never asks for them).
"""
from typing import Optional

import torch
from torch import nn

from . import _lib, linear, ops
from ._lib import check, handle, load, stream_ptr
from .nerf_base import NeRF, _NormalDot
from .nerf_helper import makeMLP
from .ref_func import generate_ide_fn
from .train_engine import PackedLinear, _dgrad_segs, _empty16, _fwd_segs, _pad8, bgrad, encode, sync_all, wgrad


class _PackedCat(PackedLinear):
    """Several nn.Linear layers reading the same input, packed as one (their rows concatenated): one GEMM for all heads."""

    def __init__(self, lins):
        self.lins = lins
        self.out_f = sum(l.weight.shape[0] for l in lins)
        self.in_f = lins[0].weight.shape[1]
        self.in_pad = _pad8(self.in_f)
        self.perm_host = self.perm = None
        self.key = None
        self.hi = self.lo = self.bias = None

    def sync(self):
        key = tuple((l.weight.data_ptr(), l.weight._version, l.bias._version) for l in self.lins)
        if key != self.key:
            w = torch.cat([l.weight.detach() for l in self.lins], dim=0).contiguous()
            b = torch.cat([l.bias.detach() for l in self.lins]).contiguous()
            if self.hi is not None and self.hi.device == w.device:      # in place: launch plans hold these pointers
                linear.to_bf16(w, ld_dst=self.in_pad, out=(self.hi, self.lo))
                self.bias.copy_(b)
            else:
                self.hi, self.lo = linear.to_bf16(w, ld_dst=self.in_pad)
                self.bias = b
            self.key = key
        return self


class RefNeRF(NeRF):
    def __init__(self, position_flevel, sh_max_level, bottle_neck_dim=128, hidden_unit=256, output_dim=256, use_srgb=False, cat_origin=True,
                 perturb_bottle_neck_w=0.1) -> None:
        super().__init__(position_flevel, cat_origin, lambda x: x)          # density is not activated during render
        self.sh_max_level = sh_max_level
        self.bottle_neck_dim = bottle_neck_dim
        self.dir_enc_dim = ((1 << sh_max_level) - 1 + sh_max_level) << 1
        extra_width = 3 if cat_origin else 0
        spatial_module_list = makeMLP(6 * position_flevel + extra_width, hidden_unit)
        for _ in range(3):
            spatial_module_list.extend(makeMLP(hidden_unit, hidden_unit))
        self.spa_block1 = nn.Sequential(*spatial_module_list)
        self.spa_block2 = nn.Sequential(
            *makeMLP(hidden_unit + 6 * position_flevel + extra_width, hidden_unit),
            *makeMLP(hidden_unit, hidden_unit), *makeMLP(hidden_unit, hidden_unit),
            *makeMLP(hidden_unit, output_dim)
        )
        self.rho_tau_head = nn.Linear(output_dim, 2)
        self.norm_col_tint_head = nn.Linear(output_dim, 9)
        self.bottle_neck = nn.Linear(output_dim, bottle_neck_dim)
        self.spec_rgb_head = nn.Sequential(*makeMLP(output_dim, 3, nn.Sigmoid()))
        dir_input_dim = 1 + bottle_neck_dim + self.dir_enc_dim
        directional_module_list = makeMLP(dir_input_dim, hidden_unit)
        for _ in range(3):
            directional_module_list.extend(makeMLP(hidden_unit, hidden_unit))
        self.dir_block1 = nn.Sequential(*directional_module_list)
        self.dir_block2 = nn.Sequential(
            *makeMLP(hidden_unit + dir_input_dim, hidden_unit),
            *makeMLP(hidden_unit, hidden_unit), *makeMLP(hidden_unit, output_dim),
            *makeMLP(hidden_unit, output_dim)
        )
        self.use_srgb = use_srgb
        self.perturb_bottle_neck_w = perturb_bottle_neck_w
        self.integrated_dir_enc = generate_ide_fn(sh_max_level)
        self.hidden_unit, self.output_dim = hidden_unit, output_dim
        self.apply(self.init_weight)
        self.precision = None      # 'bf16x3' (default) or 'bf16'

    # ---- engine ------------------------------------------------------------------------------------------------------
    def _engine(self):
        e = self.__dict__.get("_nb2_ref_engine")
        if e is None:
            if not self.cat_origin or self.hidden_unit != self.output_dim:
                raise _lib.NB2Error("RefNeRF: cat_origin=False / output_dim != hidden_unit are not supported by the engine")
            H, enc = self.hidden_unit, 3 + 6 * self.position_flevel
            din = 1 + self.bottle_neck_dim + self.dir_enc_dim
            s1, s2, d1, d2 = self.spa_block1, self.spa_block2, self.dir_block1, self.dir_block2
            e = dict(
                spa=[PackedLinear(s1[0]), PackedLinear(s1[2]), PackedLinear(s1[4]), PackedLinear(s1[6]),
                     # spa_block2.0 multiplies cat(enc, h) (ref_model.py:79); the engine's buffer is [h | enc | pad]
                     PackedLinear(s2[0], col_perm=torch.cat((torch.arange(enc) + H, torch.arange(H))), in_pad=H + _pad8(enc)),
                     PackedLinear(s2[2]), PackedLinear(s2[4]), PackedLinear(s2[6])],
                heads=_PackedCat([self.norm_col_tint_head, self.rho_tau_head]),
                bottle=PackedLinear(self.bottle_neck),
                dir=[PackedLinear(d1[0]), PackedLinear(d1[2]), PackedLinear(d1[4]), PackedLinear(d1[6]),
                     # dir_block2.0 multiplies cat(all_inputs, r) (ref_model.py:99); buffer [r | all_inputs | pad]
                     PackedLinear(d2[0], col_perm=torch.cat((torch.arange(din) + H, torch.arange(H))), in_pad=H + _pad8(din)),
                     PackedLinear(d2[2]), PackedLinear(d2[4]), PackedLinear(d2[6])],
                spec=PackedLinear(self.spec_rgb_head[0]))
            self.__dict__["_nb2_ref_engine"] = e
        return e

    max_plans = 2      # recorded batch sizes kept per module (their activation buffers stay allocated)

    def _packed(self):
        e = self._engine()
        return [*e["spa"], e["heads"], e["bottle"], *e["dir"], e["spec"]]

    def _forward_engine(self, pts2d, dirs2d, cam_dir=None, shift_softplus=False, want_grad=False):
        """pts2d (n, >=3) positions, dirs2d (n, 3) view directions -> (rgbo (n,4), normal (n,3), ndot (n,) or None[, plan]).
        The launches are recorded once per batch size as a linear.Program and replayed (one host call per run of GEMMs)."""
        x3 = self.precision != "bf16"
        if self.precision not in (None, "bf16x3", "bf16"):
            raise _lib.NB2Error(f"RefNeRF runs on the layer-wise engine: precision 'bf16x3' (default) or 'bf16', not {self.precision!r}")
        dev, n = pts2d.device, pts2d.shape[0]
        packed = self._packed()
        sync_all(packed)
        has_cam = cam_dir is not None
        plans = self.__dict__.setdefault("_nb2_ref_plans", {})
        key = (n, x3, dev, self.training, has_cam, bool(shift_softplus)) + tuple(pk.hi.data_ptr() for pk in packed)
        out = torch.empty((n, 4), dtype=torch.float32, device=dev)
        normal = torch.empty((n, 3), dtype=torch.float32, device=dev)
        ndot = torch.empty((n,), dtype=torch.float32, device=dev) if has_cam else None
        cam = _lib.f32(cam_dir).reshape(3) if has_cam else None
        dyn = dict(pts=pts2d, dirs=dirs2d, out=out, normal=normal)
        if has_cam:
            dyn.update(ndot=ndot, cam=cam)
        plan = plans.get(key)
        if plan is None or plan.busy:          # busy: its activations belong to a forward whose backward is still to come
            plans.pop(key, None)
            while len(plans) >= self.max_plans:
                plans.pop(next(iter(plans)))
            plan = plans[key] = _RefPlan()
            with linear.Program(dev) as prog:
                prog.bind(**dyn)
                plan.acts = self._record(prog, n, x3, has_cam, shift_softplus)
            plan.fwd = prog
        plan.fwd.run(**dyn)
        plan.fwd.release()
        plan.epoch += 1
        if want_grad:
            plan.busy = True
            return out, normal, ndot, plan
        return out, normal, ndot

    def _record(self, prog, n, x3, has_cam, shift_softplus):
        e = self._engine()
        dev = prog.dev
        H, enc = self.hidden_unit, 3 + 6 * self.position_flevel
        enc_w = _pad8(enc)
        din = 1 + self.bottle_neck_dim + self.dir_enc_dim
        din_w = _pad8(din)
        lib, h = load(), handle(dev)
        inp = prog.inp

        def lin(x, pk, K, out, bias, act=linear.ACT_RELU):
            linear.gemm(n, pk.out_f, _fwd_segs(x, (pk.hi, pk.lo), K, x3), bias=bias, act=act, out_hi=out[0], out_lo=out[1])

        # ---- spatial MLP (ref_model.py:70-80) ----
        C5 = _empty16(n, H + enc_w, dev, x3)
        E = (C5[0][:, H:], C5[1][:, H:] if x3 else None)
        prog.call(lambda: encode(inp["pts"], 0, self.position_flevel, False, E[0], E[1]))
        spa = e["spa"]
        bs = [p.lin.bias.detach() for p in spa]
        h1, h2, h3 = (_empty16(n, H, dev, x3) for _ in range(3))
        lin(E, spa[0], enc, h1, bs[0])
        lin(h1, spa[1], H, h2, bs[1])
        lin(h2, spa[2], H, h3, bs[2])
        lin(h3, spa[3], H, (C5[0][:, :H], C5[1][:, :H] if x3 else None), bs[3])
        h5, h6, h7, inter = (_empty16(n, H, dev, x3) for _ in range(4))
        lin(C5, spa[4], H + enc, h5, bs[4])
        lin(h5, spa[5], H, h6, bs[5])
        lin(h6, spa[6], H, h7, bs[6])
        lin(h7, spa[7], H, inter, bs[7])
        # ---- heads (ref_model.py:82-87) ----
        heads = torch.empty((n, 12), dtype=torch.float32, device=dev)
        hk = e["heads"]
        linear.gemm(n, 11, _fwd_segs(inter, (hk.hi, hk.lo), H, x3), bias=hk.bias, out_f32=heads[:, :11])
        Cd = _empty16(n, H + din_w, dev, x3)                                   # [r4 | bottleneck | ide | nv_dot | pad]
        bd = self.bottle_neck_dim
        bk = e["bottle"]
        bott = (Cd[0][:, H:H + bd], Cd[1][:, H:H + bd] if x3 else None)
        if self.training:
            # ref_model.py:86-87: Gaussian perturbation of the bottleneck (drawn by torch, applied before the bf16 split)
            b32 = torch.empty((n, bd), dtype=torch.float32, device=dev)
            linear.gemm(n, bd, _fwd_segs(inter, (bk.hi, bk.lo), H, x3), bias=self.bottle_neck.bias.detach(), out_f32=b32)

            def perturb():
                b32.add_(torch.normal(0, self.perturb_bottle_neck_w, b32.shape, device=dev))
                hi_c, lo_c = linear.to_bf16(b32, ld_dst=bd, want_lo=x3)
                bott[0].copy_(hi_c)
                if x3:
                    bott[1].copy_(lo_c)

            prog.call(perturb)
        else:
            linear.gemm(n, bd, _fwd_segs(inter, (bk.hi, bk.lo), H, x3), bias=self.bottle_neck.bias.detach(), out_hi=bott[0], out_lo=bott[1])
        # ---- geometry + integrated directional encoding (ref_model.py:88-94) ----
        reflect = torch.empty((n, 3), dtype=torch.float32, device=dev)
        rough = torch.empty((n,), dtype=torch.float32, device=dev)
        nv = torch.empty((n,), dtype=torch.float32, device=dev)
        tail = (Cd[0][:, H + bd:], Cd[1][:, H + bd:] if x3 else None)

        def geometry():
            st = stream_ptr(dev)
            dirs = inp["dirs"]
            check(lib.nb2_ref_geometry(h, heads.data_ptr(), 12, dirs.data_ptr(), dirs.stride(0), n, inp["normal"].data_ptr(), reflect.data_ptr(),
                                       rough.data_ptr(), nv.data_ptr(), st))
            ide = self.integrated_dir_enc(reflect, rough.view(n, 1))
            check(lib.nb2_ref_dir_inputs(h, ide.data_ptr(), self.dir_enc_dim, nv.data_ptr(), n, tail[0].data_ptr(), _lib.ptr_int(tail[1]),
                                         tail[0].stride(0), tail[0].shape[1], st))

        prog.call(geometry)
        # ---- directional MLP (ref_model.py:96-101) ----
        dr = e["dir"]
        bdm = [p.lin.bias.detach() for p in dr]
        A_in = (Cd[0][:, H:], Cd[1][:, H:] if x3 else None)
        r1, r2, r3 = (_empty16(n, H, dev, x3) for _ in range(3))
        lin(A_in, dr[0], din, r1, bdm[0])
        lin(r1, dr[1], H, r2, bdm[1])
        lin(r2, dr[2], H, r3, bdm[2])
        lin(r3, dr[3], H, (Cd[0][:, :H], Cd[1][:, :H] if x3 else None), bdm[3])
        q1, q2, q3, q4 = (_empty16(n, H, dev, x3) for _ in range(4))
        lin(Cd, dr[4], H + din, q1, bdm[4])
        lin(q1, dr[5], H, q2, bdm[5])
        lin(q2, dr[6], H, q3, bdm[6])
        lin(q3, dr[7], H, q4, bdm[7])
        spec = torch.empty((n, 3), dtype=torch.float32, device=dev)
        sk = e["spec"]
        linear.gemm(n, 3, _fwd_segs(q4, (sk.hi, sk.lo), H, x3), bias=self.spec_rgb_head[0].bias.detach(), act=linear.ACT_SIGMOID, out_f32=spec)
        prog.call(lambda: check(lib.nb2_ref_color(h, spec.data_ptr(), heads.data_ptr(), 12, 1 if self.use_srgb else 0, 1 if shift_softplus else 0,
                                                  inp["normal"].data_ptr(), _lib.ptr_int(inp.get("cam")), n, inp["out"].data_ptr(),
                                                  _lib.ptr_int(inp.get("ndot")), stream_ptr(dev))))
        return dict(E=E, h1=h1, h2=h2, h3=h3, C5=C5, h5=h5, h6=h6, h7=h7, inter=inter, heads=heads, Cd=Cd, A_in=A_in, r1=r1, r2=r2, r3=r3,
                    q1=q1, q2=q2, q3=q3, q4=q4, spec=spec, shift_softplus=bool(shift_softplus))

    # ---- backward plan (training) ---------------------------------------------------------------------------------
    def _grad_slots(self, flat):
        """{id(parameter): view of the flat gradient buffer}, in self.parameters() order."""
        views, off = {}, 0
        for p in self.parameters():
            views[id(p)] = flat[off:off + p.numel()].view(p.shape)
            off += p.numel()
        return views

    def _backward_engine(self, plan, pts2d, dirs2d, g_out, g_normal, x3):
        """-> (flat parameter gradients in self.parameters() order, d_pts (n, 3))."""
        dev, n = g_out.device, g_out.shape[0]
        flat = torch.empty(sum(p.numel() for p in self.parameters()), dtype=torch.float32, device=dev)
        d_pts = torch.empty((n, 3), dtype=torch.float32, device=dev)
        dyn = dict(pts=pts2d, dirs=dirs2d, g=g_out, gn=g_normal, grads=flat, d_pts=d_pts)
        if plan.bwd is None:
            with linear.Program(dev) as prog:
                prog.bind(**dyn)
                self._record_backward(prog, plan.acts, n, x3, self._grad_slots(flat))
            plan.bwd = prog
        plan.bwd.run(**dyn)
        plan.bwd.release()
        plan.busy = False
        return flat, d_pts

    def _record_backward(self, prog, a, n, x3, slot):
        e = self._engine()
        dev = prog.dev
        H, enc = self.hidden_unit, 3 + 6 * self.position_flevel
        enc_w = _pad8(enc)
        bd = self.bottle_neck_dim
        din = 1 + bd + self.dir_enc_dim
        din_w = _pad8(din)
        lib, h = load(), handle(dev)
        inp = prog.inp
        sm = torch.cuda.get_device_properties(dev).multi_processor_count
        spa, dr, hk, bk, sk = e["spa"], e["dir"], e["heads"], e["bottle"], e["spec"]
        if a["shift_softplus"]:
            raise _lib.NB2Error("RefNeRF: the density shift of the render path is applied by the caller during training (train.py:181)")

        def wg(pk, dy, x, n_out, n_in, perm=None):
            wgrad(dy, x, n_out, n_in, n, x3, slot[id(pk.lin.weight)], perm=perm, sm_count=sm, grad_b=slot[id(pk.lin.bias)])

        def dg(dy, w, K, mask, width=None):
            """dX = (dY W) [* relu mask] as bf16 hi / lo rows."""
            out = _empty16(n, H if width is None else width, dev, x3)
            linear.gemm(n, out[0].shape[1], _dgrad_segs(dy, w, K, x3), mask=mask, out_hi=out[0], out_lo=out[1])
            return out

        wl = lambda pk: (pk.hi, pk.lo)
        # ---- colour composition (ref_model.py:102-105) and the specular head ----
        ds = _empty16(n, 8, dev, x3)
        d_heads = torch.empty((n, 16), dtype=torch.float32, device=dev)
        prog.call(lambda: check(lib.nb2_ref_color_backward(h, a["spec"].data_ptr(), a["heads"].data_ptr(), 12, 1 if self.use_srgb else 0,
                                                           inp["g"].data_ptr(), n, ds[0].data_ptr(), _lib.ptr_int(ds[1]), d_heads.data_ptr(), 16,
                                                           stream_ptr(dev))))
        wg(sk, ds, a["q4"], 3, H)
        dq4 = dg(ds, wl(sk), 3, a["q4"][0])
        # ---- directional MLP (ref_model.py:96-101), last layer first ----
        wg(dr[7], dq4, a["q3"], H, H)
        dq3 = dg(dq4, wl(dr[7]), H, a["q3"][0])
        wg(dr[6], dq3, a["q2"], H, H)
        dq2 = dg(dq3, wl(dr[6]), H, a["q2"][0])
        wg(dr[5], dq2, a["q1"], H, H)
        dq1 = dg(dq2, wl(dr[5]), H, a["q1"][0])
        wg(dr[4], dq1, a["Cd"], H, H + din, perm=dr[4].perm)                       # dir_block2.0 on cat(all_inputs, r)
        dr4 = dg(dq1, dr[4].cols(0, H), H, a["Cd"][0][:, :H])
        wg(dr[3], dr4, a["r3"], H, H)
        dr3 = dg(dr4, wl(dr[3]), H, a["r3"][0])
        wg(dr[2], dr3, a["r2"], H, H)
        dr2 = dg(dr3, wl(dr[2]), H, a["r2"][0])
        wg(dr[1], dr2, a["r1"], H, H)
        dr1 = dg(dr2, wl(dr[1]), H, a["r1"][0])
        wg(dr[0], dr1, a["A_in"], H, din)
        # gradient of cat(bottleneck, ide, nv_dot): through dir_block1.0 and through the skip input of dir_block2.0
        d_in32 = torch.empty((n, din_w), dtype=torch.float32, device=dev)
        d_in = _empty16(n, din_w, dev, x3)
        linear.gemm(n, din_w, _dgrad_segs(dr1, wl(dr[0]), H, x3) + _dgrad_segs(dq1, dr[4].cols(H, H + din_w), H, x3),
                    out_f32=d_in32, out_hi=d_in[0], out_lo=d_in[1])
        db = (d_in[0][:, :bd], d_in[1][:, :bd] if x3 else None)
        wg(bk, db, a["inter"], bd, H)
        # ---- normal / reflection / IDE / roughness (ref_model.py:83-94) ----
        mat, ml = self.integrated_dir_enc.tables(dev)
        n_pow, n_pairs = mat.shape

        def geometry():
            dirs = inp["dirs"]
            check(lib.nb2_ref_geometry_backward(h, a["heads"].data_ptr(), 12, dirs.data_ptr(), dirs.stride(0), n, d_in32.data_ptr(), din_w, bd,
                                                inp["gn"].data_ptr(), mat.data_ptr(), ml.data_ptr(), n_pairs, n_pow, d_heads.data_ptr(), 16,
                                                stream_ptr(dev)))
            linear.to_bf16(d_heads, ld_dst=16, want_lo=x3, out=dh)

        dh = _empty16(n, 16, dev, x3)
        prog.keep += [mat, ml]
        prog.call(geometry)
        # ---- heads (ref_model.py:81-82; packed rows: norm_col_tint_head 0..8, rho_tau_head 9..10) ----
        nct, rt = self.norm_col_tint_head, self.rho_tau_head
        wgrad(dh, a["inter"], 11, H, n, x3, [(0, 9, slot[id(nct.weight)]), (9, 11, slot[id(rt.weight)])], sm_count=sm,
              grad_b=[(0, 9, slot[id(nct.bias)]), (9, 11, slot[id(rt.bias)])])
        d_inter = _empty16(n, H, dev, x3)
        linear.gemm(n, H, _dgrad_segs(dh, wl(hk), 11, x3) + _dgrad_segs(db, wl(bk), bd, x3), mask=a["inter"][0],
                    out_hi=d_inter[0], out_lo=d_inter[1])
        # ---- spatial MLP (ref_model.py:70-80) ----
        wg(spa[7], d_inter, a["h7"], H, H)
        dh7 = dg(d_inter, wl(spa[7]), H, a["h7"][0])
        wg(spa[6], dh7, a["h6"], H, H)
        dh6 = dg(dh7, wl(spa[6]), H, a["h6"][0])
        wg(spa[5], dh6, a["h5"], H, H)
        dh5 = dg(dh6, wl(spa[5]), H, a["h5"][0])
        wg(spa[4], dh5, a["C5"], H, H + enc, perm=spa[4].perm)                      # spa_block2.0 on cat(enc, h)
        dh4 = dg(dh5, spa[4].cols(0, H), H, a["C5"][0][:, :H])
        wg(spa[3], dh4, a["h3"], H, H)
        dh3 = dg(dh4, wl(spa[3]), H, a["h3"][0])
        wg(spa[2], dh3, a["h2"], H, H)
        dh2 = dg(dh3, wl(spa[2]), H, a["h2"][0])
        wg(spa[1], dh2, a["h1"], H, H)
        dh1 = dg(dh2, wl(spa[1]), H, a["h1"][0])
        wg(spa[0], dh1, a["E"], H, enc)
        # ---- positions: the encoding enters spa_block1.0 and the skip input of spa_block2.0 (RefNeRF.get_grad) ----
        d_enc = torch.empty((n, enc_w), dtype=torch.float32, device=dev)
        linear.gemm(n, enc_w, _dgrad_segs(dh1, wl(spa[0]), H, x3) + _dgrad_segs(dh5, spa[4].cols(H, H + enc_w), H, x3), out_f32=d_enc)

        def positions():
            pts = inp["pts"]
            check(lib.nb2_encode_backward(h, pts.data_ptr(), pts.stride(0), 0, n, self.position_flevel, d_enc.data_ptr(), enc_w,
                                          inp["d_pts"].data_ptr(), stream_ptr(dev)))

        prog.call(positions)

    def forward(self, pts: torch.Tensor, ray_d: Optional[torch.Tensor] = None):
        """pts (ray_num, point_num, 6) = [xyz, dir] (or (.., 3) with ray_d) -> ((.., 4) = [rgb, density], normal (.., 3)).
        Under autograd (a parameter or `pts` requires a gradient) the result is differentiable w.r.t. the parameters and the
        positions (train.py:176-180: fine_pos.requires_grad = True; RefNeRF.get_grad(density, fine_pos))."""
        R, P = pts.shape[0], pts.shape[1]
        wants = torch.is_grad_enabled() and (pts.requires_grad or any(p.requires_grad for p in self.parameters()))
        if torch.is_grad_enabled() and ray_d is not None and ray_d.requires_grad:
            raise _lib.NB2Error("RefNeRF.forward: gradients w.r.t. the view directions are not built (the reference never takes them)")
        if not wants:
            p2 = _lib.f32(pts.detach()).reshape(R * P, pts.shape[-1])
            d2 = p2[:, 3:6] if ray_d is None else _lib.f32(ray_d.detach()).reshape(R * P, 3)
            with torch.no_grad():
                out, normal, _ = self._forward_engine(p2, d2)
            return out.view(R, P, 4), normal.view(R, P, 3)
        p2 = _lib.f32(pts).reshape(R * P, pts.shape[-1])
        d2 = p2.detach()[:, 3:6] if ray_d is None else _lib.f32(ray_d.detach()).reshape(R * P, 3)
        out, normal = _RefFunction.apply(self, p2, d2, *self.parameters())
        return out.view(R, P, 4), normal.view(R, P, 3)

    # ---- training-side helpers of the reference (ref_model.py:107-124) ----------------------------------------------------
    @staticmethod
    def coarse_grad_select(fine_grads: torch.Tensor, sort_inds: torch.Tensor, c_pnum: int) -> torch.Tensor:
        """Rows of `fine_grads` (ray_num, all_pnum, C) that came from the coarse samples, in sorted order: the merged samples
        are cat(fine, coarse) before the sort (nerf_base.py:58-73), so the last c_pnum pre-sort positions are the coarse ones."""
        ray_num, all_pnum, _ = fine_grads.shape
        selector = torch.arange(all_pnum, device=fine_grads.device).expand(ray_num, all_pnum) >= all_pnum - c_pnum
        selector = torch.gather(selector, -1, sort_inds)
        return fine_grads[selector].reshape(ray_num, c_pnum, -1)

    @staticmethod
    def get_grad(func_val: torch.Tensor, inputs: torch.Tensor) -> torch.Tensor:
        """Normalised gradient of `func_val` (density) w.r.t. `inputs` (positions); first order, no graph (ref_model.py:118-124)."""
        grad, = torch.autograd.grad(func_val, inputs, torch.ones_like(func_val), retain_graph=True)
        grad_norm = grad.norm(dim=-1, keepdim=True)
        return grad / torch.maximum(torch.full_like(grad_norm, 1e-5), grad_norm)


class _RefPlan:
    """The recorded forward (and, once needed, backward) launches of a RefNeRF at one batch size, with the activations."""

    def __init__(self):
        self.fwd = self.bwd = None
        self.acts = None
        self.busy = False
        self.epoch = 0


class _RefFunction(torch.autograd.Function):
    """forward(module, pts2d, dirs2d, *params) -> (rgbo (n,4), normal (n,3)); params are passed so autograd routes their gradients."""

    @staticmethod
    def forward(ctx, module, pts2d, dirs2d, *params):
        pts2d = pts2d.contiguous()
        out, normal, _, plan = module._forward_engine(pts2d, dirs2d, want_grad=True)
        ctx.module, ctx.plan, ctx.epoch, ctx.x3 = module, plan, plan.epoch, module.precision != "bf16"
        ctx.pts, ctx.dirs = pts2d, dirs2d
        return out, normal

    @staticmethod
    def backward(ctx, g_out, g_normal):
        if ctx.plan.epoch != ctx.epoch:
            raise _lib.NB2Error("RefNeRF keeps one set of activations per recorded plan: the module ran forward again at this batch size "
                                "after this pass had been differentiated, so its activations are gone")
        m = ctx.module
        flat, d_pts = m._backward_engine(ctx.plan, ctx.pts, ctx.dirs, g_out.contiguous(), g_normal.contiguous(), ctx.x3)
        grads, off = [], 0
        for p in m.parameters():
            grads.append(flat[off:off + p.numel()].view(p.shape))
            off += p.numel()
        if ctx.pts.shape[1] != 3:
            full = torch.zeros_like(ctx.pts)
            full[:, :3] = d_pts
            d_pts = full
        return (None, d_pts, None, *grads)


class WeightedNormalLoss(nn.Module):
    """ref_model.py:127-135: weight (ray_num, point_num) * (1 - <d_norm, p_norm>), mean over points or sum."""

    def __init__(self, size_average=False):
        super().__init__()
        self.size_average = size_average

    def forward(self, weight: torch.Tensor, d_norm: torch.Tensor, p_norm: torch.Tensor) -> torch.Tensor:
        dot_diff = 1. - torch.sum(d_norm * p_norm, dim=-1)
        return torch.mean(weight * dot_diff) if self.size_average else torch.sum(weight * dot_diff)


class BackFaceLoss(nn.Module):
    """ref_model.py:137-143: mean of weight * relu(<normal, ray_d>) (normals facing away from the camera)."""

    def forward(self, weight: torch.Tensor, normal: torch.Tensor, ray_d: torch.Tensor) -> torch.Tensor:
        return torch.mean(weight * torch.relu(torch.sum(normal * ray_d, dim=-1)))
