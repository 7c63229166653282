#!/bin/bash
# BASELINE configs[4] image size (800x800 = 640,000 rays) on one GPU: parity spot check + throughput.
mkdir -p gpurun_out
timeout 500 python bench.py --size 800 --steps 5 --warmup 3 > gpurun_out/bench_800.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_800.log | cut -c1-300
