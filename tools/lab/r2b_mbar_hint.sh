#!/bin/bash
# Lab: A/B on one box -- mbarrier.try_wait with a suspend-time hint (libnerfb200.so) against the default time limit (libnerfb200_prev.so)
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_b_mlp.py tests/test_gpu_c_render.py tests/test_gpu_g_gemm.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -2
for rep in 1 2; do
for lib in libnerfb200.so libnerfb200_prev.so; do
  echo "$lib $(NB2_LIB=$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],2), d['roofline']['step_share'], d['clocks']['sm_mhz'], {k: round(v['ms_per_step'],2) for k,v in d['other_precisions'].items()} if d['other_precisions'] else '')")"
done
done
for lib in libnerfb200.so libnerfb200_prev.so; do echo "$lib"; NB2_LIB=$lib timeout 300 python tools/lab/r2b_gemm_time.py 2>&1 | tail -6; done
