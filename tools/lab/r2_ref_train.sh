#!/bin/bash
# Lab: Ref-NeRF training side -- backward vs autograd, the reference's training closure, plan tests.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_h_train.py tests/test_gpu_i_refnerf.py -x -q -m gpu -s -k "position or closure or plans or run_closure" > gpurun_out/ref_train_tests.log 2>&1
grep -v "Warning\|warnings.warn\|^$" gpurun_out/ref_train_tests.log | tail -40
