#!/bin/bash
# round 2: the split-precision kernel after the epilogue diet -- bench (fp16x3 only), role counters, overflow test
cd "$GRAFT_REPO_ROOT"
timeout 200 python -m pytest tests/test_gpu_b_mlp.py tests/test_gpu_d_variants.py -m gpu -q -p no:cacheprovider -k "saturate or deterministic" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2_split_bench.json 2> gpurun_out/r2_split_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2_split_bench.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'fine ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], d['roofline']['step_share'], d['clocks'])"
NB2_LIB=libnerfb200_prof.so timeout 300 python tools/gpu_probe.py roles fp16x3 > gpurun_out/r2_roles_fp16x3.txt 2>&1
NB2_LIB=libnerfb200_prof.so timeout 300 python tools/gpu_probe.py roles fp16 > gpurun_out/r2_roles_fp16.txt 2>&1
cat gpurun_out/r2_roles_fp16x3.txt | tail -18
