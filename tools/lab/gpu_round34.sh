#!/bin/bash
mkdir -p gpurun_out
timeout 200 python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/bench_quick.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_quick.log | cut -c1-160
