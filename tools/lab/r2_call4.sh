#!/bin/bash
# round 2, GPU call 4: first run of the generic tcgen05 GEMM (MN-major descriptors, cp.async producers, split-K)
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_g_gemm.py -m gpu -q -x -s -p no:cacheprovider > gpurun_out/r2c4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c4_pytest.log
tail -30 gpurun_out/r2c4_pytest.log
