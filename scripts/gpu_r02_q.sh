#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests/test_gpu_fit.py tests/test_gpu_fused.py tests/test_gpu_spans.py tests/test_gpu_large.py tests/test_gpu_basis.py -q -m gpu -x 2>&1 | tail -4
timeout 300 python scripts/prof_all.py plan bwd fit 2>&1 | tail -1
