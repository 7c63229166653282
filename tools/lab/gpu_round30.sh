#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "### $name"; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "rc=$rc $name"; tail -n ${TAILN:-3} gpurun_out/$name.log; [ $rc -eq 0 ] || exit 1; }
TMO=90 TAILN=1 run time_fp16x3 python tools/gpu_probe.py time fp16x3
TMO=90 TAILN=1 run time_fp16 python tools/gpu_probe.py time fp16
grep -h VARIANT gpurun_out/time_fp16x3.log gpurun_out/time_fp16.log
TMO=300 TAILN=3 run tests_bcd python -m pytest tests/test_gpu_b_mlp.py tests/test_gpu_c_render.py tests/test_gpu_d_variants.py -q -x --timeout=100
