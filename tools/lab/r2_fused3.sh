#!/bin/bash
# Lab: shared-tile split-precision triples in the layer-wise GEMM -- parity tests, then A/B (NB2_TC_DEBUG=8 = off) of a forward
# layer and of the training-step / Ref-NeRF legs.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_g_gemm.py tests/test_gpu_h_train.py tests/test_gpu_i_refnerf.py -x -q -m gpu 2>&1 | tail -3
for dbg in 0 8; do
  echo "== NB2_TC_DEBUG=$dbg"
  NB2_TC_DEBUG=$dbg timeout 300 python tools/lab/r2_gemm_time.py 2>&1 | tail -3
  NB2_TC_DEBUG=$dbg timeout 600 python - <<'PY' 2>&1 | grep -v Warn | tail -4
import json, torch, bench, nerf_b200
from nerf_b200 import synthetic
dev = torch.device("cuda:0")
t = bench.train_step_leg(dev)
print({k: (round(v["ms_per_step"], 3), round(v["speedup_vs_torch_cuda_fp32"], 2)) for k, v in t.items() if k.startswith("rays_")})
prop = nerf_b200.ProposalNetwork(10, 256); prop.load_state_dict(synthetic.make_params("proposal", 1, "smooth")); prop = prop.to(dev).eval()
c = bench.config4_leg(dev, prop)
print({k: round(v["ms_per_step"], 3) for k, v in c.items() if isinstance(v, dict)})
PY
done
