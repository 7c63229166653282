#!/bin/bash
# round 2: full GPU suite + smoke + default bench line + reference arm at HEAD (what the driver runs at round end)
cd "$GRAFT_REPO_ROOT"
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
tail -4 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r02_smoke.log; tail -2 gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
echo "ref rc=$?"; cut -c1-300 gpurun_out/r02_bench_reference_arm.json
