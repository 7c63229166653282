#!/bin/bash
# First run of the N-half pipelined pair kernel (NB2_TC_NHALF=1): numerics against the oracle, timing against the
# layer-serial pair kernel, role counters, then the whole GPU suite on the new kernel.
mkdir -p gpurun_out
run() { name=$1; shift; echo "### $name"; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n ${TAILN:-3} gpurun_out/$name.log; }
export NB2_TC_NHALF=1
for prec in fp16 fp16x3 bf16 bf16x3; do
  TMO=120 TAILN=4 run mlp_nerf_$prec python tools/gpu_probe.py mlp nerf $prec 5000
  TMO=120 TAILN=3 run mlp_prop_$prec python tools/gpu_probe.py mlp proposal $prec 5000
done
for prec in fp16 fp16x3; do
  NB2_TC_NHALF=0 TMO=120 TAILN=1 run time0_$prec python tools/gpu_probe.py time $prec
  NB2_TC_NHALF=1 TMO=120 TAILN=1 run time1_$prec python tools/gpu_probe.py time $prec
done
grep -h "VARIANT\|render 400" gpurun_out/time*.log
for prec in fp16x3 fp16; do
  NB2_LIB=libnerfb200_prof.so TMO=200 TAILN=16 run roles3_$prec python tools/gpu_probe.py roles $prec
done
TMO=900 run gpu_tests_nhalf python -m pytest tests -q -m gpu -x --timeout=300
TMO=300 run smoke_nhalf python __graft_entry__.py smoke
