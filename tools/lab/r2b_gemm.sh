#!/bin/bash
# round 2 (second session): GEMM tests + training / Ref-NeRF tests, then the GEMM shapes of a 256-wide layer at 1M rows
# (NB2_TC_DEBUG=256: the 8-warp epilogue for every shape)
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_g_gemm.py tests/test_gpu_h_train.py tests/test_gpu_i_refnerf.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5
for dbg in 0 256; do echo "== NB2_TC_DEBUG=$dbg"; NB2_TC_DEBUG=$dbg timeout 300 python tools/lab/r2b_gemm_time.py 2>&1 | tail -7; done
