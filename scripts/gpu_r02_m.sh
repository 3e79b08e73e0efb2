#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests/test_gpu_fit.py tests/test_gpu_contraction.py tests/test_gpu_fused.py -q -m gpu -x 2>&1 | tail -15
