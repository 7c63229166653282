#!/bin/bash
# round 2 (second session): op + render tests, then the standalone HBM-bound ops under ncu (gpu__time_duration)
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_a_ops.py tests/test_gpu_e_next_rows.py tests/test_gpu_c_render.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_hbm_launches.csv python tools/lab/r2_hbm_ops.py > gpurun_out/r2b_hbm_run.log 2>&1
python tools/lab/r2_hbm_table.py gpurun_out/r2b_hbm_launches.csv gpurun_out/r2_hbm_bytes.json 6549.1 | tee gpurun_out/r2b_hbm_table.txt
