"""Mirror of the reference's nerf/ref_model.py: Ref-NeRF (forward; SURVEY 8f-3, BASELINE configs[3]).

Same constructor, submodule names and state_dict keys as the reference (nerf/ref_model.py:16-65).  forward() runs on
the layer-wise engine: every nn.Linear is one launch of the generic tcgen05 GEMM (nb2_gemm_bf16, bf16 hi + lo operands,
fp32 accumulation), the integrated directional encoding is `ide_kernel`, and the elementwise steps between the MLPs
(normal / reflection / roughness, colour composition) are three small CUDA kernels (csrc/nb2_refnerf.cu).  The
reference's torch.cat inputs are column ranges of wider buffers (the matching weight columns are permuted once).

Not built: the training-side pieces of Ref-NeRF (density gradients w.r.t. position = double backward through the spatial
MLP, WeightedNormalLoss / BackFaceLoss): forward() is inference-only and refuses to run under autograd.
"""
from typing import Optional

import torch
from torch import nn

from . import _lib, linear, ops
from ._lib import check, handle, load, stream_ptr
from .nerf_base import NeRF, _NormalDot
from .nerf_helper import makeMLP
from .ref_func import generate_ide_fn
from .train_engine import PackedLinear, _empty16, _fwd_segs, _pad8, encode


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

    def _forward_engine(self, pts2d, dirs2d, cam_dir=None, shift_softplus=False):
        """pts2d (n, >=3) positions, dirs2d (n, 3) view directions -> (rgbo (n,4), normal (n,3), ndot (n,) or None).
        The launches are recorded once per batch size as a linear.Program and replayed (one host call per run of GEMMs)."""
        x3 = self.precision != "bf16"
        if self.precision not in (None, "bf16x3", "bf16"):
            raise _lib.NB2Error(f"RefNeRF runs on the layer-wise engine: precision 'bf16x3' (default) or 'bf16', not {self.precision!r}")
        dev, n = pts2d.device, pts2d.shape[0]
        packed = self._packed()
        for pk in packed:
            pk.sync()
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
        prog = plans.get(key)
        if prog is None:
            while len(plans) >= self.max_plans:
                plans.pop(next(iter(plans)))
            with linear.Program(dev) as prog:
                prog.bind(**dyn)
                self._record(prog, n, x3, has_cam, shift_softplus)
            plans[key] = prog
        prog.run(**dyn)
        prog.release()
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

    def forward(self, pts: torch.Tensor, ray_d: Optional[torch.Tensor] = None):
        """pts (ray_num, point_num, 6) = [xyz, dir] (or (.., 3) with ray_d) -> ((.., 4) = [rgb, density], normal (.., 3))."""
        if torch.is_grad_enabled() and (pts.requires_grad or (ray_d is not None and ray_d.requires_grad)):
            raise _lib.NB2Error("RefNeRF.forward: gradients w.r.t. positions (the reference's get_grad) are not built; run under torch.no_grad()")
        R, P = pts.shape[0], pts.shape[1]
        p2 = _lib.f32(pts.detach()).reshape(R * P, pts.shape[-1])
        d2 = p2[:, 3:6] if ray_d is None else _lib.f32(ray_d.detach()).reshape(R * P, 3)
        with torch.no_grad():
            out, normal, _ = self._forward_engine(p2, d2)
        return out.view(R, P, 4), normal.view(R, P, 3)
