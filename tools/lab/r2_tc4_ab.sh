#!/bin/bash
# Lab: A/B on one box -- chunked hand-over (default) against one hand-over per layer (NB2_TC_DEBUG=4).
mkdir -p gpurun_out
for rep in 1 2; do
for dbg in 0 4; do
  NB2_TC_DEBUG=$dbg timeout 300 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/tc4ab_$dbg.log 2>&1
  echo "debug=$dbg $(grep '"metric"' gpurun_out/tc4ab_$dbg.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['clocks'])")"
done
done
for dbg in 0 4; do
NB2_TC_DEBUG=$dbg NB2_LIB=libnerfb200_prof.so timeout 300 python tools/gpu_probe.py roles fp16x3 2>&1 | grep "mma_wait_A\|mma_total\|g0_wait_acc\|g0_epi_hidden"
done
