#!/bin/bash
# round 2 (second session), 2 GPUs: new op tests, then BASELINE configs[4] sharded over 2 GPUs + the data-parallel training step
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_f_peer.py tests/test_gpu_b_mlp.py -m gpu -q -p no:cacheprovider -k "peer or two_devices" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --train-ddp > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err
echo "bench rc=$?"; tail -5 gpurun_out/r2b_bench_n2.err; cut -c1-300 gpurun_out/r2b_bench_n2.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2b_bench_n2.json").read().strip().splitlines()[-1])
print(json.dumps(d.get("train_step_ddp"), indent=1)[:3000])
PY
