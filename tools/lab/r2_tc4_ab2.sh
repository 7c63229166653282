#!/bin/bash
# Lab: A/B on one box -- read-all-first / interleaved-chunk epilogue (libnerfb200.so) against the committed chunked one (libnerfb200_prev.so).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_b_mlp.py tests/test_gpu_c_render.py -x -q -m gpu 2>&1 | tail -2
for rep in 1 2; do
for lib in libnerfb200.so libnerfb200_prev.so; do
  NB2_LIB=$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/tc4ab2.log 2>&1
  echo "$lib $(grep '"metric"' gpurun_out/tc4ab2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['clocks'])")"
done
done
