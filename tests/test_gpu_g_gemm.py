"""GPU: the generic tcgen05 GEMM of the layer-wise engine (nb2_gemm_bf16) against torch.matmul on the same bf16 operands,
in the three operand-layout combinations the training step uses (forward, dgrad, wgrad), with K segments (split-precision
passes, torch.cat inputs), ragged sizes, N tiling, split-K and every epilogue option."""
import pytest
import torch

import nerf_b200
from nerf_b200 import linear

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(DEV)


def bf(x):
    return x.to(torch.bfloat16)


def close(got, ref, tol):
    err = float((got.float() - ref).abs().max())
    assert err <= tol * max(1.0, float(ref.abs().max())), (err, float(ref.abs().max()))


@pytest.mark.parametrize("M,K,N", [(1000, 64, 256), (128, 256, 256), (777, 320, 256), (4096, 256, 128), (130, 128, 3), (5, 64, 1), (3000, 424, 256)])
def test_forward_shape(M, K, N):
    """Y = act(X W^T + b): A K-major, B K-major; outputs fp32 + bf16 hi / lo; ragged M, padded K, tiny N."""
    X, W, b = bf(rnd((M, K), 1)), bf(rnd((N, K), 2, K ** -0.5)), rnd((N,), 3)
    ref = torch.relu(X.float() @ W.float().T + b)
    ld = (N + 7) // 8 * 8
    y32 = torch.full((M, N), -7.0, device=DEV)
    hi = torch.zeros((M, ld), dtype=torch.bfloat16, device=DEV)
    lo = torch.zeros((M, ld), dtype=torch.bfloat16, device=DEV)
    linear.gemm(M, N, [(X, False, W, False, K)], bias=b, act=linear.ACT_RELU, out_f32=y32, out_hi=hi[:, :N], out_lo=lo[:, :N])
    close(y32, ref, 2e-5)
    close(hi[:, :N].float() + lo[:, :N].float(), ref, 3e-5)
    close(hi[:, :N], ref, 8e-3)
    assert float(hi[:, N:].abs().max()) == 0.0 if ld > N else True


def test_segments_and_split_precision():
    """Three passes (lo*hi, hi*lo, hi*hi) and a concatenated input as K segments: fp32-faithful product of fp32 matrices."""
    M, K1, K2, N = 900, 256, 64, 256
    X1, X2, W = rnd((M, K1), 1), rnd((M, K2), 2), rnd((N, K1 + K2), 3, 0.06)
    x1h, x1l = linear.to_bf16(X1)
    x2h, x2l = linear.to_bf16(X2)
    wh, wl = linear.to_bf16(W)
    out = torch.empty((M, N), device=DEV)
    segs = []
    for (a1, a2, w) in ((x1l, x2l, wh), (x1h, x2h, wl), (x1h, x2h, wh)):
        segs += [(a1, False, w[:, :K1], False, K1), (a2, False, w[:, K1:], False, K2)]
    linear.gemm(M, N, segs, out_f32=out)
    ref = (torch.cat((X1, X2), -1).double() @ W.double().T).float()
    close(out, ref, 2e-5)          # bf16 hi+lo carries 16 bits: ~1e-5 relative


@pytest.mark.parametrize("M,N_out,K_in", [(1000, 256, 256), (513, 256, 320), (2000, 128, 288), (300, 16, 128)])
def test_dgrad_shape(M, N_out, K_in):
    """dX = (dY W) * (X > 0): A = dY K-major, B = W (out, in) MN-major; N = in may exceed 256 (two N tiles)."""
    dY, W, Xs = bf(rnd((M, N_out), 1)), bf(rnd((N_out, K_in), 2, 0.06)), bf(rnd((M, K_in), 3))
    ref = (dY.float() @ W.float()) * (Xs.float() > 0)
    hi = torch.empty((M, K_in), dtype=torch.bfloat16, device=DEV)
    lo = torch.empty((M, K_in), dtype=torch.bfloat16, device=DEV)
    linear.gemm(M, K_in, [(dY, False, W, True, N_out)], mask=Xs, out_hi=hi, out_lo=lo)
    close(hi.float() + lo.float(), ref, 3e-5)


@pytest.mark.parametrize("rows,N_out,K_in,splits", [(4096, 256, 256, 8), (10000, 256, 320, 37), (777, 128, 288, 3), (5000, 16, 128, 148), (64, 256, 64, 4)])
def test_wgrad_shape(rows, N_out, K_in, splits):
    """dW = dY^T X: both operands MN-major, split-K partial sums + deterministic reduction; empty splits contribute zero."""
    dY, X = bf(rnd((rows, N_out), 1)), bf(rnd((rows, K_in), 2))
    ref = dY.float().T @ X.float()
    ld_ws = (K_in + 31) // 32 * 32
    m_pad = (N_out + 127) // 128 * 128
    ws = torch.full((splits, m_pad, ld_ws), float("nan"), device=DEV)
    linear.gemm(N_out, K_in, [(dY, True, X, True, rows)], out_f32=ws.view(splits * m_pad, ld_ws)[:N_out], splits=splits, split_stride=m_pad * ld_ws)
    out = torch.empty((N_out, K_in), device=DEV)
    linear.reduce_splits(ws, splits, m_pad * ld_ws, N_out, K_in, ld_ws, out)
    close(out, ref, 1e-5 * rows ** 0.5)
    again = torch.empty_like(out)
    ws.fill_(float("nan"))
    linear.gemm(N_out, K_in, [(dY, True, X, True, rows)], out_f32=ws.view(splits * m_pad, ld_ws)[:N_out], splits=splits, split_stride=m_pad * ld_ws)
    linear.reduce_splits(ws, splits, m_pad * ld_ws, N_out, K_in, ld_ws, again)
    assert torch.equal(out, again)          # no atomics: bit-reproducible


@pytest.mark.parametrize("M,K,N", [(8192, 256, 256), (8229, 256, 128), (20000, 64, 64), (9000, 320, 512), (8197, 256, 320), (8192, 128, 32)])
@pytest.mark.parametrize("act", [linear.ACT_NONE, linear.ACT_RELU, linear.ACT_SIGMOID])
def test_bf16_output_shapes_of_the_sixteen_warp_epilogue(M, K, N, act):
    """The shapes the training step and Ref-NeRF run most: bf16 hi (+ lo) outputs only, M >= 8192 -- N = 64 / 128 / 256 / 512
    take the 16-warp epilogue (two column halves per tile, 32-column passes, 64-byte-swizzled staging tiles, TMA stores),
    N = 320 (two 160-wide tiles) and N = 32 the 8-warp one; ragged M (rows clipped by the TMA unit), split-precision passes,
    bias + every activation, hi only and hi + lo; outputs land inside a wider buffer without touching its other columns."""
    X, W, b = rnd((M, K), 1), rnd((N, K), 2, K ** -0.5), rnd((N,), 3)
    xh, xl = linear.to_bf16(X)
    wh, wl = linear.to_bf16(W)
    pre = (X.double() @ W.double().T + b.double())
    ref = {linear.ACT_NONE: pre, linear.ACT_RELU: torch.relu(pre), linear.ACT_SIGMOID: torch.sigmoid(pre)}[act].float()
    segs = [(xl, False, wh, False, K), (xh, False, wl, False, K), (xh, False, wh, False, K)]
    wide_hi = torch.full((M, N + 64), 7.0, dtype=torch.bfloat16, device=DEV)
    wide_lo = torch.full((M, N + 64), 7.0, dtype=torch.bfloat16, device=DEV)
    linear.gemm(M, N, segs, bias=b, act=act, out_hi=wide_hi[:, 32:32 + N], out_lo=wide_lo[:, 32:32 + N])
    close(wide_hi[:, 32:32 + N].float() + wide_lo[:, 32:32 + N].float(), ref, 3e-5)
    for w in (wide_hi, wide_lo):
        assert float((w[:, :32].float() - 7.0).abs().max()) == 0.0 and float((w[:, 32 + N:].float() - 7.0).abs().max()) == 0.0
    only = torch.empty((M, N), dtype=torch.bfloat16, device=DEV)
    linear.gemm(M, N, segs, bias=b, act=act, out_hi=only)
    assert torch.equal(only, wide_hi[:, 32:32 + N])


@pytest.mark.parametrize("M,N_out,K_in", [(8192, 256, 256), (8300, 256, 128), (30011, 128, 256), (8192, 256, 512)])
def test_dgrad_shape_of_the_sixteen_warp_epilogue(M, N_out, K_in):
    """dX = (dY W) * (X > 0) at M >= 8192: relu mask staged per 32-column pass and applied to the packed hi / lo words."""
    dY, W, Xs = rnd((M, N_out), 1), rnd((N_out, K_in), 2, 0.06), bf(rnd((M, K_in), 3))
    dh, dl = linear.to_bf16(dY)
    wh, wl = linear.to_bf16(W)
    ref = ((dY.double() @ W.double()) * (Xs.double() > 0)).float()
    hi = torch.empty((M, K_in), dtype=torch.bfloat16, device=DEV)
    lo = torch.empty((M, K_in), dtype=torch.bfloat16, device=DEV)
    linear.gemm(M, K_in, [(dl, False, wh, True, N_out), (dh, False, wl, True, N_out), (dh, False, wh, True, N_out)], mask=Xs, out_hi=hi, out_lo=lo)
    close(hi.float() + lo.float(), ref, 3e-5)
    assert float((hi.float() + lo.float())[Xs.float() <= 0].abs().max()) == 0.0


def test_to_bf16_permutation():
    W = rnd((256, 319), 5)
    perm = torch.cat((torch.arange(63) + 256, torch.arange(256))).to(torch.int32).to(DEV)     # [enc | hidden] -> [hidden | enc]
    hi, lo = linear.to_bf16(W, ld_dst=320, col_perm=perm)
    back = (hi.float() + lo.float())
    assert float((back[:, 256:319] - W[:, :63]).abs().max()) < 1e-4 and float((back[:, :256] - W[:, 63:]).abs().max()) < 1e-4
    assert float(back[:, 319].abs().max()) == 0.0


def test_bias_gradient_as_ones_gemm():
    from nerf_b200.train_engine import bgrad
    x = rnd((5000, 256), 6)
    xh, xl = linear.to_bf16(x)
    s = torch.empty(256, device=DEV)
    bgrad((xh, xl), 256, 5000, True, s)
    close(s, x.sum(0), 2e-4)


@pytest.mark.parametrize("rows,n_out,n_in", [(5000, 256, 256), (70000, 256, 320), (3001, 1, 256), (9000, 11, 128), (130, 128, 64), (40000, 256, 64)])
def test_bias_gradient_rides_on_the_weight_gradient(rows, n_out, n_in):
    """wgrad(..., grad_b=): the row sums of dY^T from one extra N = 16 MMA per k-step against a tile of ones
    (nb2_gemm_desc.a_rowsum_out) equal the separate ones-GEMM and the fp64 sums; the weight gradient itself is bit-identical
    with and without the extra MMAs; one and two N tiles, M below / at / above one 128-row tile, ragged row counts."""
    from nerf_b200.train_engine import bgrad, wgrad
    ld = (n_out + 7) // 8 * 8
    dy = rnd((rows, ld), 11)
    dy[:, n_out:] = 0
    x = rnd((rows, n_in), 12)
    dyp, xp = linear.to_bf16(dy), linear.to_bf16(x)
    gw0, gw1 = torch.empty(n_out, n_in, device=DEV), torch.empty(n_out, n_in, device=DEV)
    gb0, gb1 = torch.empty(n_out, device=DEV), torch.empty(n_out, device=DEV)
    wgrad(dyp, xp, n_out, n_in, rows, True, gw0)
    bgrad(dyp, n_out, rows, True, gb0)
    wgrad(dyp, xp, n_out, n_in, rows, True, gw1, grad_b=gb1)
    assert torch.equal(gw0, gw1)
    ref = dy[:, :n_out].double().sum(0)
    scale = float(dy[:, :n_out].double().abs().sum(0).max())
    assert float((gb1.double() - ref).abs().max()) <= 2e-6 * scale
    assert float((gb1 - gb0).abs().max()) <= 2e-6 * scale
    # packed heads: the rows of one GEMM reduced into two parameters
    if n_out == 11:
        a, b = torch.empty(9, device=DEV), torch.empty(2, device=DEV)
        wgrad(dyp, xp, n_out, n_in, rows, True, [(0, 9, gw1[:9]), (9, 11, gw1[9:])], grad_b=[(0, 9, a), (9, 11, b)])
        assert torch.equal(torch.cat((a, b)), gb1)


def test_launch_plan_replays_bit_identically_with_rebound_outputs():
    """linear.Program: a forward GEMM, a split-K wgrad and two deferred reductions recorded once, replayed into other output
    tensors: bit-identical to the eager launches (nb2_gemm_bf16_batch / nb2_reduce_splits_batch against the single calls)."""
    rows, N_out, K_in, splits = 3000, 256, 320, 9
    X, W, dY = bf(rnd((rows, K_in), 1)), bf(rnd((N_out, K_in), 2, 0.06)), bf(rnd((rows, N_out), 3))
    perm = torch.cat((torch.arange(64) + 256, torch.arange(256))).to(torch.int32).to(DEV)
    ld_ws, m_pad = 320, 256

    def eager():
        y = torch.empty((rows, N_out), device=DEV)
        linear.gemm(rows, N_out, [(X, False, W, False, K_in)], out_f32=y)
        outs = []
        for p in (None, perm):
            ws = torch.empty((splits, m_pad, ld_ws), device=DEV)
            linear.gemm(N_out, K_in, [(dY, True, X, True, rows)], out_f32=ws.view(splits * m_pad, ld_ws)[:N_out], splits=splits, split_stride=m_pad * ld_ws)
            g = torch.empty((N_out, K_in), device=DEV)
            linear.reduce_splits(ws, splits, m_pad * ld_ws, N_out, K_in, ld_ws, g, col_perm=p)
            outs.append(g)
        return y, outs
    y_ref, g_ref = eager()
    y = torch.empty((rows, N_out), device=DEV)
    flat = torch.empty(2 * N_out * K_in, device=DEV)
    marks = []
    with linear.Program(torch.device(DEV)) as prog:
        prog.bind(y=y, grads=flat)
        linear.gemm(rows, N_out, [(X, False, W, False, K_in)], out_f32=y)
        prog.call(lambda: marks.append(1))
        for i, p in enumerate((None, perm)):
            ws = torch.empty((splits, m_pad, ld_ws), device=DEV)
            linear.gemm(N_out, K_in, [(dY, True, X, True, rows)], out_f32=ws.view(splits * m_pad, ld_ws)[:N_out], splits=splits, split_stride=m_pad * ld_ws)
            linear.reduce_splits(ws, splits, m_pad * ld_ws, N_out, K_in, ld_ws, flat[i * N_out * K_in:(i + 1) * N_out * K_in].view(N_out, K_in), col_perm=p)
    assert not marks and float(torch.nan_to_num(y).abs().max()) >= 0       # recording launched nothing
    for _ in range(2):
        y2 = torch.full((rows, N_out), float("nan"), device=DEV)
        flat2 = torch.full((2 * N_out * K_in,), float("nan"), device=DEV)
        prog.run(y=y2, grads=flat2)
        assert torch.equal(y2, y_ref)
        assert torch.equal(flat2[:N_out * K_in].view(N_out, K_in), g_ref[0]) and torch.equal(flat2[N_out * K_in:].view(N_out, K_in), g_ref[1])
    assert len(marks) == 2
    with pytest.raises(nerf_b200.NB2Error):
        with linear.Program(torch.device(DEV)):
            linear.reduce_splits(ws, splits, m_pad * ld_ws, N_out, K_in, ld_ws, g_ref[0], accumulate=True)
    assert linear.recording() is None
