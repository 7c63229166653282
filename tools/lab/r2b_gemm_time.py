"""Lab: the three GEMM shapes of a 256-wide layer of the layer-wise engine at 1M rows (bf16x3): forward (bias + relu, hi + lo
out), dgrad (relu mask, hi + lo out), wgrad (split-K fp32 partials) and the ones-GEMM bias gradient; CUDA-event times."""
import sys, os, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "."))
from nerf_b200 import linear, train_engine
DEV = "cuda"
M, K, N = 1 << 20, 256, 256
X = torch.randn(M, K, device=DEV)
W = torch.randn(N, K, device=DEV) * 0.06
b = torch.randn(N, device=DEV)
xh, xl = linear.to_bf16(X); wh, wl = linear.to_bf16(W)
dyh, dyl = linear.to_bf16(torch.randn(M, N, device=DEV))
hi = torch.empty((M, N), dtype=torch.bfloat16, device=DEV); lo = torch.empty_like(hi)
gw = torch.empty(N, K, device=DEV); gb = torch.empty(N, device=DEV)
def fwd():
    linear.gemm(M, N, [(xl, False, wh, False, K), (xh, False, wl, False, K), (xh, False, wh, False, K)], bias=b, act=linear.ACT_RELU, out_hi=hi, out_lo=lo)
def dgrad():
    linear.gemm(M, K, [(dyl, False, wh, True, N), (dyh, False, wl, True, N), (dyh, False, wh, True, N)], mask=xh, out_hi=hi, out_lo=lo)
def wgrad():
    train_engine.wgrad((dyh, dyl), (xh, xl), N, K, M, True, gw)
def bgrad():
    train_engine.bgrad((dyh, dyl), N, M, True, gb)
def wgrad_b():
    train_engine.wgrad((dyh, dyl), (xh, xl), N, K, M, True, gw, grad_b=gb)
def fwd1():
    linear.gemm(M, N, [(xh, False, wh, False, K)], bias=b, act=linear.ACT_RELU, out_hi=hi)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
for name, fn in (("forward x3", fwd), ("dgrad x3", dgrad), ("wgrad x3", wgrad), ("bgrad x3", bgrad), ("wgrad+bias x3", wgrad_b), ("forward x1", fwd1)):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"{name:14s} {ts[2]:7.0f} us (min {ts[0]:.0f})")
# correctness of the store path against the copy-loop path is covered by tests/test_gpu_g_gemm.py
