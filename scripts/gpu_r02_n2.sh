#!/bin/bash
# two GPUs: NCCL fit / strong / c5 blocks of bench.py and the two-rank test
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_fit.py -q -m gpu -x 2>&1 | tail -15
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --inner 20 > gpurun_out/bench_n2_r02.json 2> gpurun_out/bench_n2_r02.err
echo "bench n2 rc=$?"; tail -c 5000 gpurun_out/bench_n2_r02.json; tail -n 20 gpurun_out/bench_n2_r02.err
