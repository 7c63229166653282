#!/bin/bash
# ncu launch lists at HEAD: the default bench (--no-extras) and two 8192-ray training steps
cd "$GRAFT_REPO_ROOT"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r2b_ncu_b.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_train_step_launches_8192rays.csv python tools/lab/r2_train_prof.py 8192 bf16x3 2 > gpurun_out/r2b_ncu_t.log 2>&1
wc -l gpurun_out/r02_launches_bench.csv gpurun_out/r02_train_step_launches_8192rays.csv; tail -2 gpurun_out/r2b_ncu_t.log
