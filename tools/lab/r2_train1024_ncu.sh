#!/bin/bash
# Lab: ncu launch list of the reference-default 1024-ray training step (3 steps; the last one is analysed).
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_train_step_launches_1024rays.csv python tools/lab/r2_train_prof.py 1024 bf16x3 3 > gpurun_out/train1024_ncu.log 2>&1
tail -2 gpurun_out/train1024_ncu.log
