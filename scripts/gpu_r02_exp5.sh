#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/exp5.log; : > $L
echo "== c5 count vec on/off" >> $L
timeout 300 python scripts/exp_c5.py >> $L 2>&1
THB_COUNT_VEC=0 timeout 300 python scripts/exp_c5.py >> $L 2>&1
echo "== csr variants (bulk both index widths)" >> $L
THB_CSR_BULK=2 timeout 300 python scripts/exp_csr.py >> $L 2>&1
THB_CSR_BULK=2 THB_LIB=$PWD/thb_diff_b200/_tune/lib_c768.so timeout 300 python scripts/exp_csr.py >> $L 2>&1
tail -n 20 $L
