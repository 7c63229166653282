#!/bin/bash
# Lab: role counters of the split kernel with the chunked hand-over; host profile of the 1024-ray step after the base_z cache.
mkdir -p gpurun_out
NB2_LIB=libnerfb200_prof.so timeout 300 python tools/gpu_probe.py roles fp16x3 > gpurun_out/r2_roles_fp16x3_chunks.txt 2>&1
tail -18 gpurun_out/r2_roles_fp16x3_chunks.txt
timeout 300 python tools/lab/r2_train_host.py > gpurun_out/plans_host2.log 2>&1
head -24 gpurun_out/plans_host2.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/tc4c_bench2.log 2>&1
grep '"metric"' gpurun_out/tc4c_bench2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['clocks'])"
