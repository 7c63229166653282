"""Lab: time one forward-shaped GEMM of the layer-wise engine (1M rows) under epilogue ablation flags."""
import sys, os, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "."))
from nerf_b200 import linear
DEV = "cuda"
M, K, N = 1 << 20, 256, 256
X = torch.randn(M, K, device=DEV)
W = torch.randn(N, K, device=DEV) * 0.06
b = torch.randn(N, device=DEV)
xh, xl = linear.to_bf16(X); wh, wl = linear.to_bf16(W)
hi = torch.empty((M, N), dtype=torch.bfloat16, device=DEV); lo = torch.empty_like(hi)
def run(x3, out_lo=True):
    segs = [(xl, False, wh, False, K), (xh, False, wl, False, K), (xh, False, wh, False, K)] if x3 else [(xh, False, wh, False, K)]
    linear.gemm(M, N, segs, bias=b, act=linear.ACT_RELU, out_hi=hi, out_lo=lo if out_lo else None)
for flags in (0,):
    for x3 in (True, False):
        for _ in range(2): run(x3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run(x3)
        e1.record(); torch.cuda.synchronize()
        print(f"flags {flags} x3 {x3}: {e0.elapsed_time(e1)/5*1e3:.0f} us")
run(True, out_lo=False); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): run(True, out_lo=False)
e1.record(); torch.cuda.synchronize()
print(f"x3, hi only: {e0.elapsed_time(e1)/5*1e3:.0f} us")
