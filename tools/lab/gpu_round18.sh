#!/bin/bash
# Relaxed operand-ready arrivals from the peer CTA.
mkdir -p gpurun_out
run() { name=$1; shift; echo "### $name"; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n ${TAILN:-3} gpurun_out/$name.log; }
NB2_TC_LOCKSTEP=0 TMO=120 TAILN=1 run time_fp16_ls0 python tools/gpu_probe.py time fp16
NB2_TC_LOCKSTEP=1 TMO=120 TAILN=1 run time_fp16_ls1 python tools/gpu_probe.py time fp16
TMO=120 TAILN=1 run time_fp16x3_tc2 python tools/gpu_probe.py time fp16x3
NB2_TC_TMEMA=1 TMO=120 TAILN=1 run time_fp16x3_tmema python tools/gpu_probe.py time fp16x3
grep -h "VARIANT" gpurun_out/time_*.log
NB2_TC_LOCKSTEP=0 NB2_LIB=libnerfb200_prof.so TMO=200 TAILN=14 run roles_fp16_ls0 python tools/gpu_probe.py roles fp16
NB2_TC_TMEMA=1 NB2_LIB=libnerfb200_prof.so TMO=200 TAILN=14 run roles_fp16x3_tmema python tools/gpu_probe.py roles fp16x3
NB2_TC_TMEMA=1 NB2_TC_LOCKSTEP=0 TMO=900 run gpu_tests python -m pytest tests -q -m gpu -x --timeout=300
