#!/bin/bash
# TMEM-operand kernel: where do the MMA phases spend their time?  debug 1 = no weight streaming, 2 = no MMAs.
mkdir -p gpurun_out
run() { name=$1; shift; echo "### $name"; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n ${TAILN:-3} gpurun_out/$name.log; }
export NB2_TC_TMEMA=1
for dbg in 0 1 2 3; do
  NB2_TC_DEBUG=$dbg NB2_LIB=libnerfb200_prof.so TMO=200 TAILN=14 run roles4_dbg$dbg python tools/gpu_probe.py roles fp16x3
done
