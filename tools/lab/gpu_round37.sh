#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_final.log 2>&1; echo rc=$?; grep '^{"metric"' gpurun_out/bench_n2_final.log | tail -1 | cut -c1-220
