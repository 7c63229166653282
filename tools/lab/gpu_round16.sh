#!/bin/bash
# Relaxed weight relay in every pair kernel: lockstep vs ping-pong (single pass), split, N-half; then the whole suite on the TMEM-operand kernel.
mkdir -p gpurun_out
run() { name=$1; shift; echo "### $name"; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n ${TAILN:-3} gpurun_out/$name.log; }
for ls in 1 0; do
  NB2_TC_LOCKSTEP=$ls TMO=120 TAILN=1 run time_fp16_ls$ls python tools/gpu_probe.py time fp16
done
TMO=120 TAILN=1 run time_fp16x3_tc2 python tools/gpu_probe.py time fp16x3
NB2_TC_NHALF=1 TMO=120 TAILN=1 run time_fp16_nhalf python tools/gpu_probe.py time fp16
NB2_TC_NHALF=1 TMO=120 TAILN=1 run time_fp16x3_nhalf python tools/gpu_probe.py time fp16x3
NB2_TC_TMEMA=1 TMO=120 TAILN=1 run time_fp16x3_tmema python tools/gpu_probe.py time fp16x3
grep -h "VARIANT" gpurun_out/time_*.log
NB2_TC_LOCKSTEP=0 NB2_LIB=libnerfb200_prof.so TMO=200 TAILN=14 run roles_fp16_ls0 python tools/gpu_probe.py roles fp16
NB2_TC_TMEMA=1 TMO=900 run gpu_tests_tmema python -m pytest tests -q -m gpu -x --timeout=300
