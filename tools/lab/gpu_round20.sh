#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "### $name"; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n ${TAILN:-3} gpurun_out/$name.log; }
TMO=900 TAILN=15 run gpu_tests python -m pytest tests -q -m gpu -x --timeout=300
TMO=300 run smoke python __graft_entry__.py smoke
TMO=600 TAILN=1 run bench python bench.py
