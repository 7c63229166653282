#!/bin/bash
# First run of the TMEM-operand split kernel (NB2_TC_TMEMA=1).
mkdir -p gpurun_out
run() { name=$1; shift; echo "### $name"; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n ${TAILN:-3} gpurun_out/$name.log; }
TMO=60 TAILN=2 run selftest_ts python tools/gpu_probe.py selftest_ts
export NB2_TC_TMEMA=1
for prec in fp16x3 bf16x3; do
  TMO=120 TAILN=3 run mlp_nerf_$prec python tools/gpu_probe.py mlp nerf $prec 5000
  TMO=120 TAILN=2 run mlp_prop_$prec python tools/gpu_probe.py mlp proposal $prec 5000
done
NB2_TC_TMEMA=0 TMO=120 TAILN=1 run time0_fp16x3 python tools/gpu_probe.py time fp16x3
NB2_TC_TMEMA=1 TMO=120 TAILN=1 run time1_fp16x3 python tools/gpu_probe.py time fp16x3
grep -h "VARIANT" gpurun_out/time?_fp16x3.log
NB2_LIB=libnerfb200_prof.so TMO=200 TAILN=14 run roles4_fp16x3 python tools/gpu_probe.py roles fp16x3
TMO=900 run gpu_tests_tmema python -m pytest tests -q -m gpu -x --timeout=300
