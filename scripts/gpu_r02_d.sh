#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python scripts/exp_step.py 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_spans.py tests/test_gpu_baseline_configs.py tests/test_gpu_fused.py -q -m gpu -x 2>&1 | tail -3
THB_PROFILING=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_r02d -f python scripts/prof_all.py plan > gpurun_out/ncu_r02d.log 2>&1
echo "ncu rc=$?"; tail -n 3 gpurun_out/ncu_r02d.log
