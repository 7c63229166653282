#!/bin/bash
# Lab: Ref-NeRF training tests (incl. --prop_normal) + the config-4 bench leg.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_i_refnerf.py tests/test_gpu_h_train.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python - <<'PY' > gpurun_out/ref_bench.log 2>&1
import json, torch, bench, nerf_b200
from nerf_b200 import synthetic
dev = torch.device("cuda:0")
prop = nerf_b200.ProposalNetwork(10, 256); prop.load_state_dict(synthetic.make_params("proposal", 1, "smooth")); prop = prop.to(dev).eval()
print(json.dumps(bench.config4_leg(dev, prop), indent=1))
PY
tail -40 gpurun_out/ref_bench.log
