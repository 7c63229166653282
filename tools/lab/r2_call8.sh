#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c8_bench.json 2> gpurun_out/r2c8_bench.err
echo "rc=$?"; tail -5 gpurun_out/r2c8_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c8_bench.json'))
print(json.dumps(d['train_step'], indent=1))
print(d['value'], d['e2e']['value'])
PY
