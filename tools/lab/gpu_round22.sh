#!/bin/bash
# Multi-GPU sanity: the bench contract at N=2 (engine and reference arm).
mkdir -p gpurun_out
run() { name=$1; shift; echo "### $name"; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n ${TAILN:-2} gpurun_out/$name.log; }
TMO=400 run bench_n2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3
TMO=400 run bench_ref_n2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1
