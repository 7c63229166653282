#!/bin/bash
# Lab: where a forward layer of the layer-wise GEMM spends its time (NB2_TC_DEBUG ablations; results are wrong by design).
for dbg in 0 16 32 128 144 176; do
  echo "== NB2_TC_DEBUG=$dbg  (16 no global stores, 32 no bias loads, 128 no TMEM loads)"
  NB2_TC_DEBUG=$dbg timeout 300 python tools/lab/r2_gemm_time.py 2>&1 | tail -3
done
