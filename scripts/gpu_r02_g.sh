#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_contraction.py -q -m gpu -x 2>&1 | tail -3
timeout 300 python scripts/prof_all.py csr32 csr64 2>&1 | tail -1
THB_PROFILING=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_csr -o gpurun_out/prof_r02g -f python scripts/prof_all.py csr32 > gpurun_out/ncu_r02g.log 2>&1
echo "ncu rc=$?"; tail -n 3 gpurun_out/ncu_r02g.log
