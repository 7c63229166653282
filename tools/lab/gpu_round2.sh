#!/bin/bash
# Parity suite + bench + ncu captures; logs under gpurun_out/.
mkdir -p gpurun_out
run() { name=$1; shift; echo "### $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
nvidia-smi -L
run a_ops python -m pytest tests/test_gpu_a_ops.py -q -rA --timeout=300
run b_mlp python -m pytest tests/test_gpu_b_mlp.py -q -rA --timeout=300
TAILN=25 run c_render python -m pytest tests/test_gpu_c_render.py -q -rA --timeout=600 -s
run smoke python __graft_entry__.py smoke
TAILN=3 run bench python bench.py
TAILN=3 run bench_ref python bench.py --impl reference --steps 3 --warmup 1
run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --no-extras
run ncu_full_bf16 ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 3 -c 2 -f -o gpurun_out/prof_bf16 python tools/gpu_probe.py time bf16 200
run ncu_full_fp16x3 ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 3 -c 2 -f -o gpurun_out/prof_fp16x3 python tools/gpu_probe.py time fp16x3 200
ls -la gpurun_out | head -40
