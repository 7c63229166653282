#!/bin/bash
# Lab: launch plans in the training engine -- parity of the training tests, then host time of the 1024-ray step.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_h_train.py tests/test_gpu_g_gemm.py tests/test_gpu_i_refnerf.py -x -q -m gpu > gpurun_out/plans_tests.log 2>&1
tail -5 gpurun_out/plans_tests.log
timeout 300 python tools/lab/r2_train_host.py > gpurun_out/plans_host.log 2>&1
head -60 gpurun_out/plans_host.log
