#!/bin/bash
# Round-1 evidence for the new default kernels: launch list of a bench step, ncu --set full of the MLP kernels (fp16x3: TMEM-operand
# kernel; fp16: pair kernel, ping-pong), role counters.
mkdir -p gpurun_out
run() { name=$1; shift; echo "### $name"; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n ${TAILN:-3} gpurun_out/$name.log; }
run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 2 --warmup 1 --no-extras
run ncu_full_fp16x3 ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 2 -c 2 -f -o gpurun_out/prof_fp16x3_tmema python tools/gpu_probe.py time fp16x3 400
run ncu_full_fp16 ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 2 -c 2 -f -o gpurun_out/prof_fp16_pp python tools/gpu_probe.py time fp16 400
for prec in fp16x3 fp16; do
  NB2_LIB=libnerfb200_prof.so TMO=200 TAILN=14 run roles_$prec python tools/gpu_probe.py roles $prec
done
ls -la gpurun_out | tail
