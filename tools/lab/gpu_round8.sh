#!/bin/bash
# Round-end evidence: tests, bench, launch list, ncu --set full of the MLP kernels (bench-size launch) and of the HBM-bound kernels.
mkdir -p gpurun_out
run() { name=$1; shift; echo "### $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n ${TAILN:-3} gpurun_out/$name.log; }
run gpu_tests python -m pytest tests -q -m gpu -x --timeout=600
run smoke python __graft_entry__.py smoke
TAILN=2 run bench python bench.py
TAILN=2 run bench_ref python bench.py --impl reference --steps 3 --warmup 1
run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --no-extras
run ncu_full_fp16x3 ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 2 -c 2 -f -o gpurun_out/prof_fp16x3 python tools/gpu_probe.py time fp16x3 400
run ncu_full_fp16 ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 2 -c 2 -f -o gpurun_out/prof_fp16 python tools/gpu_probe.py time fp16 400
run ncu_full_hbm ncu --set full --clock-control none -k regex:"resample_kernel|generate_rays_kernel|posenc_kernel|composite_kernel|weights_kernel|sample_pdf_kernel" -c 12 -f -o gpurun_out/prof_hbm python -m pytest tests/test_gpu_a_ops.py -q -k "positional or composite or weights_from or sample_pdf or resample"
ls -la gpurun_out | tail -20
