#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_all_shapes.py tests/test_gpu_contraction.py tests/test_gpu_baseline_configs.py -q -m gpu -x 2>&1 | tail -15
timeout 300 python scripts/prof_all.py nets eval 2>&1 | tail -1
THB_NET_PPT=2 timeout 300 python scripts/prof_all.py nets eval 2>&1 | tail -1
timeout 300 python scripts/prof_all.py csr32 csr64 2>&1 | tail -1
timeout 300 python scripts/exp_step.py 2>&1 | tail -1
