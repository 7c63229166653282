#!/bin/bash
# Lab: the whole GPU suite + smoke + bench (default line) + 1024-ray host timing.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/full_tests.log 2>&1
tail -4 gpurun_out/full_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/full_smoke.log 2>&1; tail -2 gpurun_out/full_smoke.log
timeout 300 python tools/lab/r2_train_host.py > gpurun_out/plans_host3.log 2>&1
head -3 gpurun_out/plans_host3.log
timeout 900 python bench.py > gpurun_out/full_bench.log 2>&1
grep '"metric"' gpurun_out/full_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d.get('train_step'), d.get('other_configs'))"
