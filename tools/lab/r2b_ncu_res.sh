#!/bin/bash
# ncu --set full + source counters of the fused resample kernel
cd "$GRAFT_REPO_ROOT"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"resample_kernel" --launch-skip 2 --launch-count 1 -o gpurun_out/r2b_res_full -f python tools/lab/r2_hbm_ops.py > gpurun_out/r2b_ncu_res.log 2>&1
ncu -i gpurun_out/r2b_res_full.ncu-rep --page raw --csv > gpurun_out/r2b_res_full.raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2b_res_full.raw.csv > gpurun_out/r2b_res_full.txt
ncu -i gpurun_out/r2b_res_full.ncu-rep --page source --csv > gpurun_out/r2b_res.source.csv 2>/dev/null
tail -2 gpurun_out/r2b_ncu_res.log
