#!/bin/bash
mkdir -p gpurun_out
for cfg in "1 0" "1 1" "2 1"; do set -- $cfg
  NB2_TC_CLUSTER=$1 NB2_TC_LOCKSTEP=$2 timeout 120 python tools/gpu_probe.py roles bf16 2>&1 | grep -A14 ROLES
done
NB2_TC_CLUSTER=1 timeout 120 python tools/gpu_probe.py roles fp16x3 2>&1 | grep -A14 ROLES
