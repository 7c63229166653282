#!/bin/bash
# compute-sanitizer on the default tensor kernels (small inputs): memcheck + racecheck + synccheck.
mkdir -p gpurun_out
run() { name=$1; shift; echo "### $name"; timeout ${TMO:-400} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; grep -i "ERROR SUMMARY\|RACECHECK SUMMARY\|Error:\|hazard" gpurun_out/$name.log | sort | uniq -c | head -${TAILN:-8}; tail -n 2 gpurun_out/$name.log; }
for prec in fp16x3 fp16; do
  run memcheck_$prec compute-sanitizer --tool memcheck --print-limit 5 python tools/gpu_probe.py mlp nerf $prec 700
  run racecheck_$prec compute-sanitizer --tool racecheck --print-limit 5 python tools/gpu_probe.py mlp nerf $prec 700
done
run synccheck_fp16x3 compute-sanitizer --tool synccheck --print-limit 5 python tools/gpu_probe.py mlp proposal fp16x3 700
