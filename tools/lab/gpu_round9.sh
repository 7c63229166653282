#!/bin/bash
# Re-entry check of HEAD + hardware micro-benchmarks (TMEM bandwidth, L2 weight-stream ceiling) + role counters.
mkdir -p gpurun_out
run() { name=$1; shift; echo "### $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n ${TAILN:-3} gpurun_out/$name.log; }
run gpu_tests python -m pytest tests -q -m gpu -x --timeout=600
run smoke python __graft_entry__.py smoke
TAILN=40 run microbench timeout 300 python tools/gpu_probe.py microbench
for prec in fp16x3 fp16; do
  NB2_LIB=libnerfb200_prof.so TAILN=16 run roles_$prec timeout 200 python tools/gpu_probe.py roles $prec
done
TAILN=2 run bench python bench.py
ls -la gpurun_out | tail -12
