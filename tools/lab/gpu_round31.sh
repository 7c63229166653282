#!/bin/bash
# A/B on one box: A = separate w_full / w_peer waits, B = merged stage-full barrier.
mkdir -p gpurun_out
for rep in 1 2; do
for lib in A B; do
for prec in fp16x3 fp16; do
  NB2_LIB=libnerfb200_$lib.so timeout 90 python tools/gpu_probe.py time $prec 2>&1 | grep "render 400" | tail -2 | tr '\n' ' '; echo " <- $lib $prec"
done; done; done
