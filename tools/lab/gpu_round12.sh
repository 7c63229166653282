#!/bin/bash
# Layer-serial pair kernel with the encodings moved into the MMA wait windows.
mkdir -p gpurun_out
run() { name=$1; shift; echo "### $name"; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n ${TAILN:-3} gpurun_out/$name.log; }
for prec in fp16 fp16x3; do
  TMO=120 TAILN=1 run time_$prec python tools/gpu_probe.py time $prec
done
grep -h "VARIANT" gpurun_out/time_*.log
TMO=900 run gpu_tests python -m pytest tests -q -m gpu -x --timeout=300
for prec in fp16x3 fp16; do
  NB2_LIB=libnerfb200_prof.so TMO=200 TAILN=16 run roles_$prec python tools/gpu_probe.py roles $prec
done
