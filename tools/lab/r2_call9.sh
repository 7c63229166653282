#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python -m pytest tests/test_gpu_e_next_rows.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -v "^$" | tail -12
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c9_bench.json 2> gpurun_out/r2c9_bench.err
echo "rc=$?"; tail -3 gpurun_out/r2c9_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c9_bench.json'))
print(json.dumps(d['other_configs'], indent=1))
PY
