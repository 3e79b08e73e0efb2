#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests/test_gpu_grid.py -q -m gpu -x 2>&1 | tail -3
timeout 300 python scripts/exp_grid2.py 2>&1 | tail -1
