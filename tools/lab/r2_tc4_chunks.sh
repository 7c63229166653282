#!/bin/bash
# Lab: chunked hand-over in the split kernel -- parity (render + mlp tests), then the bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_b_mlp.py tests/test_gpu_c_render.py -x -q -m gpu > gpurun_out/tc4c_tests.log 2>&1
tail -5 gpurun_out/tc4c_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/tc4c_bench.log 2>&1
grep '"metric"' gpurun_out/tc4c_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','roofline')})"
