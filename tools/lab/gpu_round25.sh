#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_probe.py hbm > gpurun_out/hbm_ops.log 2>&1; echo rc=$?; grep "HBM" gpurun_out/hbm_ops.log; tail -3 gpurun_out/hbm_ops.log
