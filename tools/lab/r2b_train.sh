#!/bin/bash
# round 2 (second session): training-side tests, then the default bench line (train_step / config-4 legs)
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_g_gemm.py tests/test_gpu_h_train.py tests/test_gpu_i_refnerf.py tests/test_gpu_b_mlp.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | cut -c1-200
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2b_bench_n1.json").read().strip().splitlines()[-1])
t = d["train_step"]
print({k: (round(v["ms_per_step"], 3), round(v["roofline"]["frac"], 3), round(v["speedup_vs_torch_cuda_fp32"], 2)) for k, v in t.items() if k.startswith("rays")})
c = d["other_configs"]["config4_refnerf"]
print({k: round(v["ms_per_step"], 3) for k, v in c.items() if isinstance(v, dict)})
print(d["value"], d["roofline"]["frac"])
PY
