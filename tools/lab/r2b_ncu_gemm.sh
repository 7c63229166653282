#!/bin/bash
# ncu --set full + source counters of one forward and one dgrad launch of the layer-wise GEMM (1M rows, 256 x 256, bf16x3)
cd "$GRAFT_REPO_ROOT"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/r2b_gemm_fwd -f python tools/lab/r2b_gemm_time.py > gpurun_out/r2b_ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel --launch-skip 10 --launch-count 1 -o gpurun_out/r2b_gemm_dgrad -f python tools/lab/r2b_gemm_time.py >> gpurun_out/r2b_ncu_gemm.log 2>&1
for f in r2b_gemm_fwd r2b_gemm_dgrad; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/$f.raw.csv > gpurun_out/$f.txt
  ncu -i gpurun_out/$f.ncu-rep --page source --csv > gpurun_out/$f.source.csv 2>/dev/null
done
grep -E "time_duration" gpurun_out/r2b_gemm_fwd.txt gpurun_out/r2b_gemm_dgrad.txt
