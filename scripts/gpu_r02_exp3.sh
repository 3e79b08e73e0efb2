#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/exp3.log; : > $L
timeout 600 python -m pytest tests/test_gpu_contraction.py -q -m gpu -x --tb=short >> $L 2>&1; echo "tests rc=$?" >> $L
timeout 600 python scripts/exp_csr_fwd.py >> $L 2>&1
for v in c1728w8 c3072w4; do THB_LIB=$PWD/thb_diff_b200/_tune/lib_$v.so timeout 600 python scripts/exp_csr_fwd.py >> $L 2>&1; done
tail -n 30 $L
