#!/bin/bash
# round 2, GPU call 1: full GPU test-suite after the hygiene / parity-theorem changes, smoke, default bench line
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2c1_gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2c1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c1_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c1_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2c1_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
echo "bench rc=$?" >> gpurun_out/r2c1_bench.err
tail -5 gpurun_out/r2c1_pytest.log; tail -3 gpurun_out/r2c1_smoke.log; cut -c1-400 gpurun_out/r2c1_bench.json
