#!/bin/bash
# round 2 (second session), 8 GPUs: BASELINE configs[4] sharded 8-way + the data-parallel training step
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 --train-ddp > gpurun_out/r2b_bench_n8.json 2> gpurun_out/r2b_bench_n8.err
echo "bench rc=$?"; tail -3 gpurun_out/r2b_bench_n8.err | cut -c1-200
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2b_bench_n8.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("shard_invariance_max_abs_diff"))
print(json.dumps(d.get("train_step_ddp"))[:1500])
PY
