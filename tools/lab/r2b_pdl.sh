#!/bin/bash
# Lab: programmatic dependent launch of the layer-wise GEMMs -- tests, then A/B of the training / Ref-NeRF legs (NB2_TC_DEBUG=512 = plain launches)
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_g_gemm.py tests/test_gpu_h_train.py tests/test_gpu_i_refnerf.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
for dbg in 0 512 0 512; do
NB2_TC_DEBUG=$dbg timeout 900 python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
t = d['train_step']; c = d['other_configs']['config4_refnerf']
print('dbg $dbg', {k: round(v['ms_per_step'], 3) for k, v in t.items() if k.startswith('rays')}, {k: round(v['ms_per_step'], 3) for k, v in c.items() if isinstance(v, dict)})
"
done
