#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_all_shapes.py tests/test_gpu_c5.py tests/test_gpu_large.py tests/test_gpu_baseline_configs.py -q -m gpu -x 2>&1 | tail -15
timeout 300 python scripts/exp_step.py 2>&1 | tail -1
THB_NET_MONO=0 timeout 300 python scripts/exp_step.py 2>&1 | tail -1
timeout 300 python scripts/prof_all.py nets eval 2>&1 | tail -1
THB_NET_MONO=0 timeout 300 python scripts/prof_all.py nets eval 2>&1 | tail -1
THB_PROFILING=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_r02e -f python scripts/prof_all.py nets eval > gpurun_out/ncu_r02e.log 2>&1
echo "ncu rc=$?"; tail -n 3 gpurun_out/ncu_r02e.log
