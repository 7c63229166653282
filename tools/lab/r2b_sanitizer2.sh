#!/bin/bash
# round 2 (second session, late): compute-sanitizer on the GEMM kernels at HEAD -- 16-warp epilogue (64-byte-swizzled staging tiles,
# TMA stores, packed relu mask), bias gradient on the wgrad launch (ones tile in the staging area, extra TMEM columns), reused B tiles,
# programmatic dependent launch -- and on the tiled encoder / Ref-NeRF glue
cd "$GRAFT_REPO_ROOT"
SEL='test_bf16_output_shapes_of_the_sixteen_warp_epilogue and (8229 or 9000 or 8197) and not 2] or test_dgrad_shape_of_the_sixteen_warp_epilogue and (8300 or 30011) or test_bias_gradient_rides and (5000 or 70000 or 3001 or 9000) or test_launch_plan or test_mlp_backward_vs_autograd or test_refnerf_forward_vs_reference'
timeout 1800 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 7 python -m pytest tests/test_gpu_g_gemm.py tests/test_gpu_h_train.py tests/test_gpu_i_refnerf.py -m gpu -q -p no:cacheprovider -k "$SEL" > gpurun_out/r02_sanitizer_memcheck_gemm_head.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck_gemm_head.log
tail -5 gpurun_out/r02_sanitizer_memcheck_gemm_head.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_g_gemm.py -m gpu -q -p no:cacheprovider -k "test_bf16_output_shapes_of_the_sixteen_warp_epilogue and 8229 and 1] or test_dgrad_shape_of_the_sixteen_warp_epilogue and 8300 or test_bias_gradient_rides and 5000" > gpurun_out/r02_sanitizer_racecheck_gemm_head.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck_gemm_head.log
tail -6 gpurun_out/r02_sanitizer_racecheck_gemm_head.log
