#!/bin/bash
# Lab: compositing deferred to the next tile's layer-2 window (libnerfb200.so) against in-place compositing (libnerfb200_prev.so)
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_b_mlp.py tests/test_gpu_c_render.py tests/test_gpu_d_variants.py tests/test_gpu_e_next_rows.py -q -m gpu -p no:cacheprovider 2>&1 | tail -2 | cut -c1-120
NB2_LIB=libnerfb200_prof.so timeout 300 python tools/gpu_probe.py roles fp16x3 2>&1 | grep -E "mma_wait_A|mma_wait_W|mma_total|g0_epi_last|g0_encode|by layer"
for rep in 1 2; do
for lib in libnerfb200.so libnerfb200_prev.so; do
  echo "$lib $(NB2_LIB=$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],2), d['roofline']['step_share'], d['clocks']['sm_mhz'])")"
done
done
