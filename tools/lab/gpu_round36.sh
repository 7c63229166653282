#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "### $name"; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "rc=$rc $name"; tail -n ${TAILN:-3} gpurun_out/$name.log; [ $rc -eq 0 ] || exit 1; }
TMO=200 TAILN=4 run tests_all python -m pytest tests -q -m gpu -x --timeout=100
TMO=120 TAILN=14 run hbm python tools/gpu_probe.py hbm
