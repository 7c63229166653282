#!/bin/bash
# round 2, GPU call 3: the tests that failed in call 1 after the theorem thresholds were set from measurements; smoke; bench line
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_c_render.py tests/test_gpu_a_ops.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2c3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c3_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c3_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2c3_smoke.log
tail -4 gpurun_out/r2c3_pytest.log; tail -2 gpurun_out/r2c3_smoke.log
