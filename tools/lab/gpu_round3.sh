#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "### $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n ${TAILN:-4} gpurun_out/$name.log; }
nvidia-smi -L
# correctness of the default variant first
run b_mlp python -m pytest tests/test_gpu_b_mlp.py -q -x --timeout=300
run c_render python -m pytest tests/test_gpu_c_render.py -q -x --timeout=600
# variant sweep
for c in 1 2 4; do for l in 0 1; do
  TAILN=1 NB2_TC_CLUSTER=$c NB2_TC_LOCKSTEP=$l run sweep_bf16_c${c}_l${l} timeout 120 python tools/gpu_probe.py time bf16
done; done
for c in 1 2 4; do
  TAILN=1 NB2_TC_CLUSTER=$c run sweep_x3_c${c} timeout 120 python tools/gpu_probe.py time fp16x3
done
grep -h VARIANT gpurun_out/sweep_*.log
