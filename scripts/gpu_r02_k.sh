#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m pytest tests/test_gpu_contraction.py -q -m gpu 2>&1 | tail -2
for cfg in "THB_BWD_PERM=-1 THB_BWD_VEC=1" "THB_BWD_PERM=-1 THB_BWD_VEC=0" "THB_BWD_PERM=1000003 THB_BWD_VEC=1" "THB_BWD_PERM=1000003 THB_BWD_VEC=0" "THB_BWD_PERM=0 THB_BWD_VEC=0" "THB_BWD_PERM=104729 THB_BWD_VEC=0"; do
  echo "$cfg"; env $cfg timeout 300 python scripts/prof_all.py csr32 csr64 2>&1 | tail -1
done
