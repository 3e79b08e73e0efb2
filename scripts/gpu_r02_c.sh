#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for c in 16 8 6 4; do THB_PLAN_CTAS_PER_SM=$c timeout 300 python scripts/exp_step.py 2>&1 | tail -1; done
THB_LUT_EXACT=0 timeout 300 python scripts/exp_step.py 2>&1 | tail -1
bash scripts/gpu_tests_all.sh
