#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_g_gemm.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
for dbg in 0 64; do echo "== NB2_TC_DEBUG=$dbg"; NB2_TC_DEBUG=$dbg timeout 300 python tools/lab/r2b_gemm_time.py 2>&1 | tail -6; done
