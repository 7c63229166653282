#!/bin/bash
# round 2: compute-sanitizer on the kernels added this round (layer-wise GEMM with TMA + staging tile, backward ray ops, Ref-NeRF glue,
# IPE producer, peer-less render) -- memcheck, then racecheck on the GEMM's shared-memory staging
cd "$GRAFT_REPO_ROOT"
SEL='test_forward_shape and (130-128-3 or 5-64-1 or 777-320) or test_wgrad_shape and 64-256-64 or test_dgrad_shape and 300-16 or test_ray_op_backward or test_refnerf_forward_vs_reference or test_coarse_fine_merge_index or test_ipe_fused and fp16x3 or test_bias_gradient'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_g_gemm.py tests/test_gpu_h_train.py tests/test_gpu_i_refnerf.py tests/test_gpu_e_next_rows.py -m gpu -q -p no:cacheprovider -k "$SEL" > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck.log
tail -6 gpurun_out/r02_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_g_gemm.py -m gpu -q -p no:cacheprovider -k "test_forward_shape and 130-128-3 or test_dgrad_shape and 300-16 or test_wgrad_shape and 64-256-64" > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck.log
tail -8 gpurun_out/r02_sanitizer_racecheck.log
