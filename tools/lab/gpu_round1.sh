#!/bin/bash
# First bring-up battery (each step in its own process, with its own timeout); logs under gpurun_out/.
mkdir -p gpurun_out
run() { name=$1; shift; echo "### $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n 6 gpurun_out/$name.log; }
nvidia-smi -L
run a_ops python -m pytest tests/test_gpu_a_ops.py -q -rA --timeout=300
run probe_selftest python tools/gpu_probe.py selftest
run probe_simt_prop python tools/gpu_probe.py mlp proposal fp32
run probe_simt_nerf python tools/gpu_probe.py mlp nerf fp32
run probe_tc_prop_bf16 timeout 120 python tools/gpu_probe.py mlp proposal bf16
run probe_tc_nerf_bf16 timeout 120 python tools/gpu_probe.py mlp nerf bf16
run probe_tc_prop_x3 timeout 120 python tools/gpu_probe.py mlp proposal bf16x3
run probe_tc_nerf_x3 timeout 120 python tools/gpu_probe.py mlp nerf bf16x3
run b_mlp python -m pytest tests/test_gpu_b_mlp.py -q -rA --timeout=300
run c_render python -m pytest tests/test_gpu_c_render.py -q -rA --timeout=600 -s
run time_fp32 python tools/gpu_probe.py time fp32 200
run time_x3 timeout 120 python tools/gpu_probe.py time bf16x3
run time_bf16 timeout 120 python tools/gpu_probe.py time bf16
