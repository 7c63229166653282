#!/bin/bash
# Strong-scaling mode of the bench (BASELINE configs[4]: one 800x800 image sharded across the GPUs), N = 1 and N = 2.
mkdir -p gpurun_out
timeout 200 python bench.py --scaling strong --size 800 --steps 5 --warmup 3 --no-extras > gpurun_out/bench_strong_n1.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_strong_n1.log | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --scaling strong --size 800 --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_strong_n2.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_strong_n2.log | cut -c1-200
