#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/exp6.log; : > $L
for v in c640n2 c768n2; do echo "== $v" >> $L; THB_CSR_BULK=2 THB_LIB=$PWD/thb_diff_b200/_tune/lib_$v.so timeout 300 python scripts/exp_csr.py >> $L 2>&1; done
tail -n 20 $L
