#!/bin/bash
# TMEM-operand kernel with four epilogue warpgroups (640 threads).  Stops at the first failing stage.
mkdir -p gpurun_out
run() { name=$1; shift; echo "### $name"; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "rc=$rc $name"; tail -n ${TAILN:-3} gpurun_out/$name.log; [ $rc -eq 0 ] || exit 1; }
export NB2_TC_GROUPS=4
TMO=60 TAILN=3 run mlp_nerf_fp16x3 python tools/gpu_probe.py mlp nerf fp16x3 5000
TMO=60 TAILN=2 run mlp_prop_fp16x3 python tools/gpu_probe.py mlp proposal fp16x3 5000
TMO=90 TAILN=1 run time_g4 python tools/gpu_probe.py time fp16x3
grep -h "VARIANT" gpurun_out/time_g4.log
