#!/bin/bash
# ncu launch list of the Ref-NeRF render at 512 rays (config 4) and 10,000 rays
cd "$GRAFT_REPO_ROOT"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_refnerf_render_launches_512rays.csv python tools/lab/r2b_ref_probe.py 16 32 > gpurun_out/r2b_ref.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_refnerf_render_launches_10000rays.csv python tools/lab/r2b_ref_probe.py 100 100 >> gpurun_out/r2b_ref.log 2>&1
tail -3 gpurun_out/r2b_ref.log; wc -l gpurun_out/r02_refnerf_render_launches_*.csv
