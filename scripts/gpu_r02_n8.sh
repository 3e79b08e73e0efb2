#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in ${NS:-8 4}; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 --inner 40 > gpurun_out/bench_n${n}_r02.json 2> gpurun_out/bench_n${n}_r02.err
echo "bench n$n rc=$?"; tail -c 300 gpurun_out/bench_n${n}_r02.json; tail -n 5 gpurun_out/bench_n${n}_r02.err | cut -c1-300
done
