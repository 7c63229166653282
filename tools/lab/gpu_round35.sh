#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "### $name"; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "rc=$rc $name"; tail -n ${TAILN:-3} gpurun_out/$name.log; [ $rc -eq 0 ] || exit 1; }
TMO=200 TAILN=6 run tests_a_c python -m pytest tests/test_gpu_a_ops.py tests/test_gpu_c_render.py -q -x --timeout=100
TMO=120 TAILN=1 run bench_quick python bench.py --steps 10 --warmup 3 --no-extras
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.log').read().strip().splitlines()[-1])
print("STEP", d['ms_per_step'], d['value'], d['roofline']['step_share'])
PY
