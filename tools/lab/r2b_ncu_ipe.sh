#!/bin/bash
# ncu --set full of the standalone IPE and posenc kernels (why they sit below the copy bandwidth)
cd "$GRAFT_REPO_ROOT"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ipe_kernel|posenc3_kernel|weights_regs|composite_regs" --launch-skip 8 --launch-count 4 -o gpurun_out/r2b_hbm_full -f python tools/lab/r2_hbm_ops.py > gpurun_out/r2b_ncu_ipe.log 2>&1
ncu -i gpurun_out/r2b_hbm_full.ncu-rep --page raw --csv > gpurun_out/r2b_hbm_full.raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2b_hbm_full.raw.csv > gpurun_out/r2b_hbm_full.txt
ncu -i gpurun_out/r2b_hbm_full.ncu-rep --page source --csv -k regex:ipe_kernel > gpurun_out/r2b_ipe.source.csv 2>/dev/null
tail -3 gpurun_out/r2b_ncu_ipe.log; wc -l gpurun_out/r2b_hbm_full.txt
