#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_a_ops.py -m gpu -q -p no:cacheprovider 2>&1 | tail -30
