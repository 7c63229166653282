#!/bin/bash
# New defaults: single pass = pair kernel, ping-pong tiles + shared epilogue; split = TMEM-operand kernel.
mkdir -p gpurun_out
run() { name=$1; shift; echo "### $name"; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n ${TAILN:-3} gpurun_out/$name.log; }
for prec in fp16 bf16 fp16x3; do
  TMO=120 TAILN=1 run time_$prec python tools/gpu_probe.py time $prec
done
grep -h "VARIANT" gpurun_out/time_*.log
TMO=900 TAILN=15 run gpu_tests python -m pytest tests -q -m gpu -x --timeout=300
NB2_LIB=libnerfb200_prof.so TMO=200 TAILN=14 run roles_fp16 python tools/gpu_probe.py roles fp16
TMO=300 run smoke python __graft_entry__.py smoke
TMO=600 TAILN=1 run bench python bench.py
