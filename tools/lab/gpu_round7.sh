#!/bin/bash
for ls in 1 0; do
  NB2_TC_LOCKSTEP=$ls timeout 120 python tools/gpu_probe.py roles bf16 2>&1 | grep -A14 ROLES
done
timeout 120 python tools/gpu_probe.py roles fp16x3 2>&1 | grep -A14 ROLES
