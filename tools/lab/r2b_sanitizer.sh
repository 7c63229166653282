#!/bin/bash
# round 2 (second session): compute-sanitizer on what this session changed -- the register-form scan kernels, the dims-3 encoder,
# slab-staged stores, the lockstep resample, the GEMM's TMA-store epilogue and packed relu mask -- memcheck, then racecheck
# (--report-api-errors no: the runtime's own lazy-loading probe, cuKernelGetFunction -> CUDA_ERROR_INVALID_HANDLE on the first launch of the
#  process, is otherwise counted as an error; it is not a memory access)
cd "$GRAFT_REPO_ROOT"
SEL='test_scan_kernels_across_sample_counts or test_posenc_and_length2pts or test_resample_fused_equals_staged or test_sample_coarse_bit_exact or test_ipe or test_forward_shape and (130-128-3 or 777-320) or test_dgrad_shape or test_segments_and_split or test_launch_plan'
timeout 1500 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 7 python -m pytest tests/test_gpu_a_ops.py tests/test_gpu_g_gemm.py -m gpu -q -p no:cacheprovider -k "$SEL" > gpurun_out/r02_sanitizer_memcheck_session2.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck_session2.log
tail -5 gpurun_out/r02_sanitizer_memcheck_session2.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_a_ops.py tests/test_gpu_g_gemm.py -m gpu -q -p no:cacheprovider -k "test_scan_kernels_across_sample_counts and (64 or 192) or test_posenc_and_length2pts or test_resample_fused_equals_staged or test_dgrad_shape and 300-16 or test_forward_shape and 130-128-3" > gpurun_out/r02_sanitizer_racecheck_session2.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck_session2.log
tail -8 gpurun_out/r02_sanitizer_racecheck_session2.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_hbm_launches.csv python tools/lab/r2_hbm_ops.py > gpurun_out/r2b_hbm_run.log 2>&1
python tools/lab/r2_hbm_table.py gpurun_out/r2b_hbm_launches.csv gpurun_out/r2_hbm_bytes.json 6549.1 | tee gpurun_out/r2b_hbm_table.txt
