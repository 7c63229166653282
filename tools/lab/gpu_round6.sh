#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "### $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n ${TAILN:-3} gpurun_out/$name.log; }
run b_mlp python -m pytest tests/test_gpu_b_mlp.py -q -x --timeout=120
run c_render python -m pytest tests/test_gpu_c_render.py -q -x --timeout=300
for ls in 0 1; do
  TAILN=1 NB2_TC_LOCKSTEP=$ls run time_pair_bf16_ls$ls timeout 120 python tools/gpu_probe.py time bf16
done
TAILN=1 run time_pair_fp16x3 timeout 120 python tools/gpu_probe.py time fp16x3
TAILN=1 run time_pair_fp16 timeout 120 python tools/gpu_probe.py time fp16
grep -h VARIANT gpurun_out/time_*.log
