#!/bin/bash
# Full validation of the committed defaults: GPU tests, smoke, bench (engine + reference arm).  Stops early on failure.
mkdir -p gpurun_out
run() { name=$1; shift; echo "### $name"; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "rc=$rc $name"; tail -n ${TAILN:-3} gpurun_out/$name.log; [ $rc -eq 0 ] || exit 1; }
TMO=500 TAILN=4 run gpu_tests python -m pytest tests -q -m gpu -x --timeout=120
TMO=200 run smoke python __graft_entry__.py smoke
TMO=400 TAILN=1 run bench python bench.py
TMO=200 TAILN=1 run bench_ref python bench.py --impl reference --steps 3 --warmup 1
