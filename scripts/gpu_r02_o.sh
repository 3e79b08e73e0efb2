#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests/test_gpu_c5.py tests/test_gpu_fused.py tests/test_gpu_all_shapes.py tests/test_gpu_large.py tests/test_gpu_baseline_configs.py -q -m gpu -x 2>&1 | tail -8
timeout 600 python scripts/exp_c5.py 2>&1 | tail -1
THB_NET_NO_GRAD4=1 timeout 600 python scripts/exp_c5.py 2>&1 | tail -1
timeout 300 python scripts/prof_all.py nets eval 2>&1 | tail -1
