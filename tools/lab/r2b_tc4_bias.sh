#!/bin/bash
# Lab: split fused kernel without the unused Wh stage of bias-only chunks -- parity tests, role counters, bench line
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_b_mlp.py tests/test_gpu_c_render.py tests/test_gpu_d_variants.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
NB2_LIB=libnerfb200_prof.so timeout 300 python tools/gpu_probe.py roles fp16x3 > gpurun_out/r2b_roles_fp16x3.txt 2>&1
grep -E "mma_wait_A|mma_wait_W|mma_total|g0_wait_acc|g0_epi" gpurun_out/r2b_roles_fp16x3.txt | tail -6
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['step_share'], d['clocks']['sm_mhz'])"
done
