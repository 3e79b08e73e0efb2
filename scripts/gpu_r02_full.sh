#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
bash scripts/gpu_tests_all.sh 2>&1 | tail -40
timeout 1200 python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_r02.json; tail -n 8 gpurun_out/bench_r02.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_r02.json 2> gpurun_out/bench_ref_r02.err; echo "ref rc=$?"; tail -c 800 gpurun_out/bench_ref_r02.json
