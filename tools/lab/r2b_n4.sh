#!/bin/bash
# round 2 (second session), 4 GPUs: BASELINE configs[4] sharded 4-way + the data-parallel training step
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 5 --train-ddp > gpurun_out/r2b_bench_n4.json 2> gpurun_out/r2b_bench_n4.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2b_bench_n4.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("shard_invariance_max_abs_diff"))
print({k: (round(v["ms_per_step"], 3), round(v["rays_per_s"]), v["replica_param_max_abs_diff"]) for k, v in d["train_step_ddp"].items() if isinstance(v, dict)})
PY
