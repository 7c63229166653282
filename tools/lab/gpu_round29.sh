#!/bin/bash
# Weak-scaling check at N=8 (the driver's own scaling run uses the same command line).
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 --no-extras > gpurun_out/bench_n8.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_n8.log | cut -c1-400
