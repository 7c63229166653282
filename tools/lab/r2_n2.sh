#!/bin/bash
# round 2: two GPUs -- peer-store gather test, strong-scaling bench line (config 5), NCCL comparison
cd "$GRAFT_REPO_ROOT"
nvidia-smi topo -m > gpurun_out/r2_n2_topo.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_f_peer.py tests/test_gpu_b_mlp.py -m gpu -q -p no:cacheprovider -k "peer or two_devices" > gpurun_out/r2_n2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_n2_pytest.log
tail -5 gpurun_out/r2_n2_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_n2_bench.json 2> gpurun_out/r2_n2_bench.err
echo "bench rc=$?"; tail -5 gpurun_out/r2_n2_bench.err; cut -c1-1500 gpurun_out/r2_n2_bench.json
