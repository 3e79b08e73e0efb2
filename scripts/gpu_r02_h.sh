#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
ls thb_diff_b200/_variants/ 2>&1 | head
timeout 600 python -m pytest tests/test_gpu_contraction.py tests/test_gpu_fused.py -q -m gpu -x 2>&1 | tail -3
timeout 300 python scripts/prof_all.py csr32 csr64 2>&1 | tail -1
timeout 300 python scripts/prof_all.py nets eval 2>&1 | tail -1
for v in m12 m20 m24; do echo $v; THB_LIB=thb_diff_b200/_variants/lib_$v.so timeout 300 python scripts/prof_all.py eval 2>&1 | tail -1; done
