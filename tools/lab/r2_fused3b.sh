#!/bin/bash
# Lab: tile sharing restricted to the wgrad shape -- tests, then A/B of the training legs (NB2_TC_DEBUG=8 = off).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_g_gemm.py tests/test_gpu_h_train.py tests/test_gpu_i_refnerf.py -q -m gpu 2>&1 | tail -3
for dbg in 0 8 0 8; do
  echo "== NB2_TC_DEBUG=$dbg"
  NB2_TC_DEBUG=$dbg timeout 600 python - <<'PY' 2>&1 | grep -v Warn | tail -2
import json, torch, bench, nerf_b200
dev = torch.device("cuda:0")
t = bench.train_step_leg(dev)
print({k: (round(v["ms_per_step"], 3), round(v["speedup_vs_torch_cuda_fp32"], 2)) for k, v in t.items() if k.startswith("rays_")})
PY
done
