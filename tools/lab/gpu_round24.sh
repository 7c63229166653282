#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "### $name"; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n ${TAILN:-3} gpurun_out/$name.log; }
TMO=120 TAILN=1 run time_fp16x3 python tools/gpu_probe.py time fp16x3
grep -h "VARIANT" gpurun_out/time_fp16x3.log
NB2_LIB=libnerfb200_prof.so TMO=200 TAILN=14 run roles_fp16x3 python tools/gpu_probe.py roles fp16x3
TMO=900 TAILN=4 run gpu_tests python -m pytest tests -q -m gpu -x --timeout=300
