#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "### $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n ${TAILN:-4} gpurun_out/$name.log; }
run probe_pair_bf16 timeout 60 python tools/gpu_probe.py mlp nerf bf16 1000
run probe_pair_x3 timeout 60 python tools/gpu_probe.py mlp nerf fp16x3 1000
run b_mlp python -m pytest tests/test_gpu_b_mlp.py -q -x --timeout=120
run c_render python -m pytest tests/test_gpu_c_render.py -q -x --timeout=300
for prec in bf16 fp16x3; do
  TAILN=1 run time_pair_$prec timeout 120 python tools/gpu_probe.py time $prec
  TAILN=1 NB2_TC_PAIR=0 NB2_TC_CLUSTER=1 NB2_TC_LOCKSTEP=1 run time_single_$prec timeout 120 python tools/gpu_probe.py time $prec
done
TAILN=14 run roles_pair_bf16 timeout 120 python tools/gpu_probe.py roles bf16
TAILN=14 run roles_pair_x3 timeout 120 python tools/gpu_probe.py roles fp16x3
grep -h VARIANT gpurun_out/time_*.log
