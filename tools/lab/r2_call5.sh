#!/bin/bash
# round 2, GPU call 5: first run of the training step (layer-wise engine forward / dgrad / wgrad, ray-op backward kernels)
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_h_train.py tests/test_gpu_b_mlp.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2c5_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c5_pytest.log
grep -n "^E  \|FAILED\|passed\|failed\|worst\|loss" gpurun_out/r2c5_pytest.log | cut -c1-300 | head -60
