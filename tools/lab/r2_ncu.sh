#!/bin/bash
# round 2: ncu evidence at HEAD -- launch list of the default bench, --set full of the dominant kernels (fine mlp_tc4 fp16x3,
# fine mlp_tc2 fp16, one hidden-layer forward launch of the layer-wise GEMM inside a training step)
cd "$GRAFT_REPO_ROOT"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r2_ncu_b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mlp_tc4_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/r02_tc4_fp16x3_fine -f python tools/gpu_probe.py time fp16x3 > gpurun_out/r2_ncu_tc4.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:mlp_tc2_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/r02_tc2_fp16_fine -f python tools/gpu_probe.py time fp16 > gpurun_out/r2_ncu_tc2.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:gemm_bf16_kernel --launch-skip 67 --launch-count 1 -o gpurun_out/r02_gemm_fwd -f python tools/lab/r2_train_prof.py 8192 bf16x3 2 > gpurun_out/r2_ncu_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep
for f in r02_tc4_fp16x3_fine r02_tc2_fp16_fine r02_gemm_fwd; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
done
ncu -i gpurun_out/r02_tc4_fp16x3_fine.ncu-rep --page source --csv > gpurun_out/r02_tc4_fp16x3_fine.source.csv 2>/dev/null
ls -la gpurun_out/*.csv | tail
