#!/bin/bash
# ncu --set full + source counters of the fine launch of the split fused kernel (fp16x3), 400x400
cd "$GRAFT_REPO_ROOT"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_tc4_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/r2b_tc4_fine -f python tools/gpu_probe.py time fp16x3 > gpurun_out/r2b_ncu_tc4.log 2>&1
ncu -i gpurun_out/r2b_tc4_fine.ncu-rep --page raw --csv > gpurun_out/r2b_tc4_fine.raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2b_tc4_fine.raw.csv > gpurun_out/r2b_tc4_fine.txt
ncu -i gpurun_out/r2b_tc4_fine.ncu-rep --page source --csv > gpurun_out/r2b_tc4_fine.source.csv 2>/dev/null
grep -E "time_duration|tensor_cycles_active" gpurun_out/r2b_tc4_fine.txt; tail -2 gpurun_out/r2b_ncu_tc4.log
