#!/bin/bash
# round 2: eight GPUs -- BASELINE configs[4]: one 800x800 image sharded 8-way, fused peer-store gather vs NCCL all_gather
cd "$GRAFT_REPO_ROOT"
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_n${N}_bench.json 2> gpurun_out/r2_n${N}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_n${N}_bench.err; cut -c1-200 gpurun_out/r2_n${N}_bench.json
