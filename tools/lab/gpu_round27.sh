#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_e_next_rows.py tests/test_gpu_d_variants.py -q -x --timeout=100 > gpurun_out/tests_e_d.log 2>&1; echo rc=$?; tail -15 gpurun_out/tests_e_d.log
